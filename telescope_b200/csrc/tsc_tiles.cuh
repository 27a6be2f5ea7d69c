// Fused E-step + M-step accumulation over flat 128-entry tiles -- the per-iteration hot kernel.
//
// Measured design inputs (B200, tools/microbench, profiles/):
//   * one (sub-)warp per read tops out at ~85 G entries/s before any scatter-add; a flat coalesced stream with a
//     K-vector gather runs at ~350 G entries/s  -> the entry stream is processed flat, not read by read;
//   * the scatter-add is the wall: RED.ADD.F64 costs one L1TEX/L2 request per 32-byte SECTOR a warp instruction
//     touches (134 G/s with 32 scattered lanes, 410 G/s when the 32 lanes fall into 8 sectors)
//     -> consecutive lanes take consecutive entries, so a read's neighbouring loci share sectors, and the caller's
//        locus numbering (neighbouring loci of a family are adjacent) is kept on the device.
//
// The entry stream is cut into tiles of whole reads, at most 128 consecutive entries each.  One warp takes one
// tile at a time, in 4 rounds of 32 entries: lane l owns entries base + 32e + l, e = 0..3.
//   n = Q * (pi*theta)[locus]                         coalesced 256 B / 128 B loads, gather through L1/L2
//   row sums: per round a Kogge-Stone segmented scan (5 shuffle steps) driven by the round's 32-bit head-flag word
//             (precomputed per tile); the open read's partial sum is carried from round to round
//   the lane holding a read's last entry publishes the read's total to a per-warp shared scratch; one lane per
//   read then computes g = w*Y / total (one fp64 division sequence per tile instead of one per entry)
//   c = n * g is scatter-added (RED.ADD.F64) into one of R accumulator replicas resident in L2.
//
// Per entry this moves 12 B of HBM (8 B Q + 4 B locus) plus 32 B of tile descriptor per ~120 entries and 8 B of
// w*Y per read; z is never written.  Reads longer than a tile take the long-row path (two passes over that read).
//
// Reference semantics: model.py:718-722 (E-step) + model.py:730-733 (M-step sums); unique reads (Y=0) have
// w*Y = 0 and add nothing, as in the reference where they enter pi only through pisum0 (model.py:699,738).
#pragma once
#include "tsc_kernels.cuh"

namespace tsc {

struct __align__(16) Tile {
    long long base;        // first entry of the tile (= first entry of its first read)
    int row0;              // first read of the tile (shard-local index)
    int meta;              // bits 0..7: end (entries in the tile, 1..128; 0 = long-read tile), bits 16..23: reads
    unsigned flags[4];     // bit p: entry base+p starts a read; plus a sentinel bit at `end` when end < 128.
                           // long-read tile: flags[0..1] = read length (64-bit)
};
static_assert(sizeof(Tile) == 32, "tile descriptor is one 32-byte sector");

constexpr int kTileWarps = 16;                 // warps per block of the tile kernel
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kScratch = 132;                  // doubles of per-warp scratch (128 reads + slack)

template <bool SMEM_TAB>
__device__ __forceinline__ double gather_pt(const double* __restrict__ pt, const double* s_tab, int s_cols, int c) {
    if (SMEM_TAB) return (c < s_cols) ? s_tab[c] : __ldg(pt + c);
    return __ldg(pt + c);
}

// streaming loads: read once, keep them out of L1 so the pi*theta table stays resident there
__device__ __forceinline__ double ld_stream(const double* p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

template <bool SMEM_TAB>
__global__ void __launch_bounds__(kTileThreads)
k_fused_tiles(const Tile* __restrict__ tiles, long long n_tiles, const double* __restrict__ q, const int* __restrict__ col,
              const double* __restrict__ wy, const double* __restrict__ pt, double* __restrict__ acc, int K, int R,
              int s_cols, const EmState* __restrict__ st) {
    extern __shared__ double s_dyn[];
    if (st->done) return;
    double* s_scr = s_dyn + (threadIdx.x >> 5) * kScratch;       // per-warp scratch
    const double* s_tab = s_dyn + kTileWarps * kScratch;          // optional copy of pt[0 .. s_cols)
    if (SMEM_TAB) {
        double* t = s_dyn + kTileWarps * kScratch;
        for (int i = threadIdx.x; i < s_cols; i += blockDim.x) t[i] = pt[i];
        __syncthreads();
    }
    double* my = acc + (size_t)(blockIdx.x % R) * K;
    const int lane = threadIdx.x & 31;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);          // lanes <= mine
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    for (; t < n_tiles; t += nwarps) {
        const int4 d0 = __ldg(reinterpret_cast<const int4*>(tiles + t));
        const uint4 fl = __ldg(reinterpret_cast<const uint4*>(tiles + t) + 1);
        const long long base = ((long long)(unsigned)d0.x) | ((long long)d0.y << 32);
        const int row0 = d0.z;
        const int end = d0.w & 0xff, nrows = (d0.w >> 16) & 0xff;

        if (end == 0) {
            // ---- long read: the whole warp walks it twice
            const long long len = ((long long)fl.x) | ((long long)fl.y << 32);
            const long long hi = base + len;
            double sum = 0;
            for (long long p = base + lane; p < hi; p += 32)
                sum += ld_stream(q + p) * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, ld_stream(col + p));
            sum = group_sum<32>(sum, 0xffffffffu);
            const double g = wy[row0] * recip0(sum);
            if (g != 0.0) {
                for (long long p = base + lane; p < hi; p += 32) {
                    const int cc = col[p];
                    const double c = (q[p] * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc)) * g;
                    if (c != 0.0) atomicAdd(my + cc, c);
                }
            }
            continue;
        }

        // ---- regular tile: 4 rounds of 32 consecutive entries
        const unsigned F[4] = {fl.x, fl.y, fl.z, fl.w};
        int cc[4];
        double qq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // all 8 loads in flight before anything is consumed (entries past the
            cc[e] = ld_stream(col + base + 32 * e + lane);   // tile are padding or the next tile's: harmless)
            qq[e] = ld_stream(q + base + 32 * e + lane);
        }
        double n[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const bool valid = (32 * e + lane) < end;
            n[e] = valid ? qq[e] * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc[e]) : 0.0;
        }
        // independent segmented inclusive scans of the four rounds
        double x[4];
        int lr[4];
        int below = 0;                     // read starts before the current round
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned hl = F[e] & le_mask;
            const int headlane = hl ? (31 - __clz(hl)) : 0;
            double v = n[e];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, v, d);
                if (lane - d >= headlane) v += y;
            }
            x[e] = v;
            lr[e] = below + __popc(hl) - 1;
            below += __popc(F[e]);
        }
        // carry the open read's partial sum from round to round and publish read totals
        double carry = 0.0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((F[e] & le_mask) == 0u) x[e] += carry;            // still inside the read that was open at round start
            // an entry closes its read when the next entry starts one (position 128 counts as a start)
            const unsigned nxt = (e < 3) ? F[e + 1] : 1u;
            const unsigned closes = (F[e] >> 1) | (nxt << 31);
            const bool last = (closes >> lane) & 1u;
            if (last && (32 * e + lane) < end) s_scr[lr[e]] = x[e];
            const double x31 = __shfl_sync(0xffffffffu, x[e], 31);
            carry = (closes >> 31) ? 0.0 : x31;
        }
        __syncwarp();
        // one lane per read: g = (w*Y) * recip0(total)
        for (int rho = lane; rho < nrows; rho += 32) {
            const double w = wy[row0 + rho];
            s_scr[rho] = (w != 0.0) ? w * recip0(s_scr[rho]) : 0.0;
        }
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((32 * e + lane) < end) {
                const double c = n[e] * s_scr[lr[e]];
                if (c != 0.0) atomicAdd(my + cc[e], c);
            }
        }
        __syncwarp();   // scratch is reused by the next tile
    }
}

}  // namespace tsc
