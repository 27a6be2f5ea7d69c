// Fused E-step + M-step accumulation over flat 128-entry tiles -- the per-iteration hot kernel.
//
// Why not one (sub-)warp per read: measured on B200 (tools/microbench, 1e8 entries, K = 30k) the row-per-subwarp
// schemes top out at ~85 G entries/s before any scatter-add, while a flat 128-bit stream with a K-vector gather
// runs at ~350 G entries/s.  So the entry stream is cut into tiles of whole reads that span at most 128 consecutive
// entries from a 4-aligned base; one warp takes one tile at a time:
//
//   lane l owns entries base+4l .. base+4l+3   -> one 128-bit load of locus indices, two of Q values (coalesced,
//                                                 512 B + 1 KiB contiguous per warp)
//   n = Q * (pi*theta)[locus]                     (gather from L1/L2, or from a shared-memory copy of the table)
//   row sums: serial over the lane's 4 entries + Kogge-Stone segmented scan of the lane tails (5 shuffle steps),
//             using the tile's precomputed 128-bit head-flag mask (all row bookkeeping is bit arithmetic on it)
//   the lane holding a read's last entry publishes the read's total to a per-warp shared scratch; one lane per
//   read then computes g = w*Y / total (one fp64 division sequence per tile, not per entry)
//   c = n * g is scatter-added (RED.ADD.F64) into one of R accumulator replicas resident in L2.
//
// Per entry this moves 12 B of HBM (8 B Q + 4 B locus) plus 32 B of tile descriptor per ~115 entries and 8 B of
// w*Y per read; z is never written.  Reads longer than a tile (> ~125 entries) take the long-row path (two passes
// over that read only).
//
// Reference semantics: model.py:718-722 (E-step) + model.py:730-733 (M-step sums); unique reads (Y=0) have
// w*Y = 0 and add nothing, as in the reference where they enter pi only through pisum0 (model.py:699,738).
#pragma once
#include "tsc_kernels.cuh"

namespace tsc {

struct __align__(16) Tile {
    long long base;        // first entry covered; multiple of 4
    int row0;              // first read of the tile (shard-local index)
    int meta;              // bits 0..7: end (valid entries are [first, end), 0 = long-row tile),
                           // bits 8..9: first, bits 16..23: number of reads
    unsigned flags[4];     // bit p: entry base+p starts a read; plus a sentinel bit at `end` when end < 128.
                           // long-row tile: flags[0..1] = row length (64-bit)
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
    const int wsel = lane >> 3, sh = (lane & 7) * 4;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    for (; t < n_tiles; t += nwarps) {
        const int4 d0 = __ldg(reinterpret_cast<const int4*>(tiles + t));
        const uint4 fl = __ldg(reinterpret_cast<const uint4*>(tiles + t) + 1);
        const long long base = ((long long)(unsigned)d0.x) | ((long long)d0.y << 32);
        const int row0 = d0.z;
        const int end = d0.w & 0xff, first = (d0.w >> 8) & 3, nrows = (d0.w >> 16) & 0xff;

        if (end == 0) {
            // ---- long read: the whole warp walks it twice
            const long long len = ((long long)fl.x) | ((long long)fl.y << 32);
            const long long lo = base + first, hi = lo + len;
            double sum = 0;
            for (long long p = base + 4 * lane; p < hi; p += 128) {
                const int4 c4 = *reinterpret_cast<const int4*>(col + p);
                const double2 qa = *reinterpret_cast<const double2*>(q + p);
                const double2 qb = *reinterpret_cast<const double2*>(q + p + 2);
                if (p + 0 >= lo && p + 0 < hi) sum += qa.x * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, c4.x);
                if (p + 1 >= lo && p + 1 < hi) sum += qa.y * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, c4.y);
                if (p + 2 >= lo && p + 2 < hi) sum += qb.x * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, c4.z);
                if (p + 3 >= lo && p + 3 < hi) sum += qb.y * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, c4.w);
            }
            sum = group_sum<32>(sum, 0xffffffffu);
            const double g = wy[row0] * recip0(sum);
            if (g != 0.0) {
                for (long long p = base + 4 * lane; p < hi; p += 128) {
                    const int4 c4 = *reinterpret_cast<const int4*>(col + p);
                    const double2 qa = *reinterpret_cast<const double2*>(q + p);
                    const double2 qb = *reinterpret_cast<const double2*>(q + p + 2);
                    const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
                    const double qq[4] = {qa.x, qa.y, qb.x, qb.y};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (p + e >= lo && p + e < hi) {
                            const double c = (qq[e] * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc[e])) * g;
                            if (c != 0.0) atomicAdd(my + cc[e], c);
                        }
                    }
                }
            }
            continue;
        }

        // ---- regular tile
        const long long p0 = base + 4 * lane;
        const int4 c4 = *reinterpret_cast<const int4*>(col + p0);
        const double2 qa = *reinterpret_cast<const double2*>(q + p0);
        const double2 qb = *reinterpret_cast<const double2*>(q + p0 + 2);
        const int cc[4] = {c4.x, c4.y, c4.z, c4.w};
        const double qq[4] = {qa.x, qa.y, qb.x, qb.y};

        // five flag bits: my 4 entries + the entry after them (position 128 counts as a read start)
        const unsigned w_lo = wsel == 0 ? fl.x : wsel == 1 ? fl.y : wsel == 2 ? fl.z : fl.w;
        const unsigned w_hi = wsel == 0 ? fl.y : wsel == 1 ? fl.z : wsel == 2 ? fl.w : 1u;
        const unsigned five = __funnelshift_r(w_lo, w_hi, sh) & 0x1fu;
        // read starts strictly before my first entry
        const int pc0 = __popc(fl.x), pc1 = pc0 + __popc(fl.y), pc2 = pc1 + __popc(fl.z);
        const int below = (wsel == 0 ? 0 : wsel == 1 ? pc0 : wsel == 2 ? pc1 : pc2) + __popc(w_lo & ((1u << sh) - 1u));

        double n[4];
        bool valid[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int p = 4 * lane + e;
            valid[e] = (p >= first) && (p < end);
            n[e] = valid[e] ? qq[e] * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc[e]) : 0.0;
        }

        // tail = sum of my entries from my last read start on (all four if none starts here)
        const unsigned f4 = five & 0xfu;
        double tail = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) { if ((f4 >> e) & 1u) tail = 0; tail += n[e]; }
        // segmented inclusive scan of the tails across lanes; a lane with a read start begins a new segment
        const unsigned heads = __ballot_sync(0xffffffffu, f4 != 0u);
        const unsigned le = heads & (0xffffffffu >> (31 - lane));
        const int headlane = le ? (31 - __clz(le)) : 0;
        double x = tail;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane - d >= headlane) x += y;
        }
        double carry = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) carry = 0.0;

        // publish read totals: an entry followed by a read start (or the tile end) closes its read
        double run = carry;
        int lr[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((five >> e) & 1u) run = 0.0;
            run += n[e];
            lr[e] = below + __popc(five & ((2u << e) - 1u)) - 1;
            if (valid[e] && ((five >> (e + 1)) & 1u)) s_scr[lr[e]] = run;
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
            if (valid[e]) {
                const double c = n[e] * s_scr[lr[e]];
                if (c != 0.0) atomicAdd(my + cc[e], c);
            }
        }
        __syncwarp();   // scratch is reused by the next tile
    }
}

}  // namespace tsc
