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
//   row sums: the numerators are transposed through a per-warp shared scratch into a blocked layout (lane l holds
//             entries 4l..4l+3), summed serially per lane and combined by ONE Kogge-Stone segmented scan (5 predicated
//             shuffle steps) driven by the tile's precomputed 128-bit head-flag mask
//   the lane holding a read's last entry publishes the read's total to a per-warp shared scratch; one lane per
//   read then computes g = w*Y / total (one fp64 division sequence per tile instead of one per entry)
//   c = n * g is scatter-added (RED.ADD.F64) into one of R accumulator replicas resident in L2.
//
// Per entry this moves 12 B of HBM (8 B Q + 4 B locus) plus 32 B of tile descriptor per ~120 entries and 8 B of
// w*Y per read; z is never written.  Reads longer than a tile get a tile of their own: up to 256 entries stay in
// registers (8 per lane, one pass), longer ones are walked twice.
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

// Greedy tiling step shared by host and device: the tile that starts at read r; returns the first read after it.
__host__ __device__ inline long long next_tile(const long long* ip, long long r, long long r_stop, Tile& t) {
    const long long base = ip[r];
    t.base = base;
    t.row0 = (int)r;
    t.flags[0] = t.flags[1] = t.flags[2] = t.flags[3] = 0;
    if (ip[r + 1] - base > 128) {   // long read: a tile of its own
        const long long len = ip[r + 1] - base;
        t.meta = (1 << 16);
        t.flags[0] = (unsigned)(len & 0xffffffffLL);
        t.flags[1] = (unsigned)(len >> 32);
        return r + 1;
    }
    long long r2 = r;
    while (r2 < r_stop && ip[r2 + 1] - base <= 128) {
        const int p = (int)(ip[r2] - base);
        t.flags[p >> 5] |= 1u << (p & 31);
        ++r2;
    }
    const int end = (int)(ip[r2] - base);
    if (end < 128) t.flags[end >> 5] |= 1u << (end & 31);
    t.meta = (end & 0xff) | ((int)(r2 - r) << 16);
    return r2;
}

// Tiling on the device: reads are cut into chunks of kChunkRows; every chunk is tiled greedily by one thread (a tile
// never crosses a chunk boundary, which costs at most one short tile per chunk).  Pass 1 counts, pass 2 writes.
constexpr int kChunkRows = 2048;

__global__ void k_tile_count(const long long* __restrict__ ip, long long n_rows, int* __restrict__ counts, int n_chunks) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    long long r = (long long)c * kChunkRows;
    const long long stop = min(r + (long long)kChunkRows, n_rows);
    int n = 0;
    Tile t;
    while (r < stop) { r = next_tile(ip, r, stop, t); ++n; }
    counts[c] = n;
}

__global__ void k_tile_fill(const long long* __restrict__ ip, long long n_rows, const long long* __restrict__ offsets,
                            int n_chunks, Tile* __restrict__ out, unsigned long long* __restrict__ n_long) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chunks) return;
    long long r = (long long)c * kChunkRows;
    const long long stop = min(r + (long long)kChunkRows, n_rows);
    long long o = offsets[c];
    unsigned long long nl = 0;
    while (r < stop) {
        Tile t;
        r = next_tile(ip, r, stop, t);
        if ((t.meta & 0xff) == 0) ++nl;
        out[o++] = t;
    }
    if (nl) atomicAdd(n_long, nl);
}

// Read pointers as the caller has them (int32 or int64, absolute) -> int64 relative to the shard's first entry,
// with validation: flags |= 1 if not non-decreasing, |= 2 if some read is empty.
template <typename T>
__global__ void k_indptr_prepare(const T* __restrict__ in, long long n_plus1, long long first, long long* __restrict__ out,
                                 int* __restrict__ flags) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    int f = 0;
    for (; i < n_plus1; i += stride) {
        const long long v = (long long)in[i];
        out[i] = v - first;
        if (i + 1 < n_plus1) {
            const long long nx = (long long)in[i + 1];
            if (nx < v) f |= 1;
            if (nx == v) f |= 2;
        }
    }
    if (f) atomicOr(flags, f);
}

// Block shape of the tile kernel.  Measured on B200 (profiles/r1_occupancy_sweep.md): 32 warps per SM at <= 64
// registers per thread is the sweet spot; a 1024-thread block pins the compiler to that budget for every epilogue.
#ifndef TSC_TILE_WARPS
#define TSC_TILE_WARPS 32
#endif
#ifndef TSC_TILE_MINBLOCKS
#define TSC_TILE_MINBLOCKS 1
#endif
constexpr int kTileWarps = TSC_TILE_WARPS;     // warps per block of the tile kernel
constexpr int kTileThreads = kTileWarps * 32;
constexpr int kScrN = 136;                     // per-warp scratch: the tile's numerators (swizzled layout transposition)
constexpr int kScrG = 132;                     //                   per-read totals, then per-read scale g
constexpr int kScratch = kScrN + kScrG;        // doubles per warp

#ifndef TSC_TILE_PREFETCH
#define TSC_TILE_PREFETCH 1
#endif
constexpr bool TILE_PREFETCH = TSC_TILE_PREFETCH != 0;
// [p, p + bytes) -> L2, as one TMA bulk prefetch (SASS UBLKPF.L2): 16-byte granules, so the range is widened to them;
// the entry arrays carry 256 entries of padding behind the last tile
__device__ __forceinline__ void tile_prefetch_l2(const void* p, unsigned bytes) {
    const unsigned long long a = reinterpret_cast<unsigned long long>(p);
    const unsigned long long a0 = a & ~15ULL;
    const unsigned n = (unsigned)((a + bytes + 15ULL - a0) & ~15ULL);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a0), "r"(n) : "memory");
}

template <bool SMEM_TAB>
__device__ __forceinline__ double gather_pt(const double* __restrict__ pt, const double* s_tab, int s_cols, int c) {
    if (SMEM_TAB) return (c < s_cols) ? s_tab[c] : __ldg(pt + c);
    return __ldg(pt + c);
}

// entry stream loads (read-only path; an L1::no_allocate hint measured neutral, profiles/r1_kernel_variants.md)
__device__ __forceinline__ double ld_stream(const double* p) { return __ldg(p); }
__device__ __forceinline__ int ld_stream(const int* p) { return __ldg(p); }
// v += y when a >= b (predicated add: one DADD instead of two selects and an add)
__device__ __forceinline__ void add_if_ge(double& v, double y, int a, int b) {
    asm("{\n\t.reg .pred p;\n\tsetp.ge.s32 p, %2, %3;\n\t@p add.rn.f64 %0, %0, %1;\n\t}" : "+d"(v) : "d"(y), "r"(a), "r"(b));
}

// What a tile pass produces.
//   TILE_FUSED : E-step + M-step sums -- c = n * (w*Y/total) scatter-added into acc (the per-iteration kernel)
//   TILE_Z     : E-step alone         -- z = n * recip0(total) written per entry (model.py:702-722; self.z; Q.norm(1))
//   TILE_LNL   : log-likelihood       -- sum z * log1p(Q * inner[locus]) with z from the E-step tables (model.py:744-760)
enum { TILE_FUSED = 0, TILE_Z = 1, TILE_LNL = 2 };

struct TileArgs {
    const Tile* tiles;
    long long n_tiles;
    const double* q;
    const int* col;
    const double* wy;          // FUSED
    const double* tab_amb;     // pi*theta of the E-step (ambiguous reads)
    const double* tab_uni;     // pi of the E-step (unique reads; Z and LNL only)
    double* acc;               // FUSED: R replicas of K doubles
    int K, R, s_cols;
    const EmState* st;         // nullptr = always run
    double* z_out;             // Z
    const double* inner_amb;   // LNL: pi*theta and pi inside log1p
    const double* inner_uni;
    double* partials;          // LNL: one partial sum per block
    const LogTab* log_tab;     // LNL: table of log1p_big
};

// LONG8: compile the single-pass path for reads of 129..256 entries (8 per lane in registers).  It wins big when
// such reads are common (Zipf rows: 140 -> 192 EM iterations/s) and costs ~2 % when there are none (its registers
// shape the allocation of the regular loop), so the host instantiates it only for shards that have long reads.
template <int MODE, bool SMEM_TAB, bool LONG8>
__global__ void __launch_bounds__(kTileThreads, TSC_TILE_MINBLOCKS)
k_tiles(const TileArgs a) {
    extern __shared__ double s_dyn[];
    __shared__ double s_red[32];
    __shared__ LogTab s_log[MODE == TILE_LNL ? kLogTab : 1];
    if (a.st && a.st->done) return;
    if (MODE == TILE_LNL) {
        for (int i = threadIdx.x; i < kLogTab; i += blockDim.x) s_log[i] = a.log_tab[i];
        __syncthreads();
    }
    const Tile* __restrict__ tiles = a.tiles;
    const double* __restrict__ q = a.q;
    const int* __restrict__ col = a.col;
    const double* __restrict__ pt = a.tab_amb;
    const int s_cols = a.s_cols;
    const long long n_tiles = a.n_tiles;
    double* s_n = s_dyn + (threadIdx.x >> 5) * kScratch;         // per-warp scratch
    double* s_g = s_n + kScrN;
    const double* s_tab = s_dyn + kTileWarps * kScratch;          // optional copy of pt[0 .. s_cols)
    if (SMEM_TAB) {
        double* t = s_dyn + kTileWarps * kScratch;
        for (int i = threadIdx.x; i < s_cols; i += blockDim.x) t[i] = pt[i];
        __syncthreads();
    }
    double* my = (MODE == TILE_FUSED) ? a.acc + (size_t)(blockIdx.x % a.R) * a.K : nullptr;
    double lnl_local = 0.0;
    const int lane = threadIdx.x & 31;
    const unsigned le_mask = 0xffffffffu >> (31 - lane);          // lanes <= mine
    const int wsel = lane >> 3, sh = (lane & 7) * 4;              // blocked layout: my 4 flag bits live in F[wsel] >> sh
    // Scratch swizzle for the transposition: entry p lives at double index 2*P + (p&1), P = (c>>1) + 36*(c&1), c = p>>1.
    // Writes (lane-consecutive, p = 32e + lane) and the two 128-bit reads per lane (entries 4l..4l+1 at 2l, entries
    // 4l+2..4l+3 at 72 + 2l) are then both free of bank conflicts.
    const int wl = 2 * (lane >> 2) + 72 * ((lane >> 1) & 1) + (lane & 1);
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    long long t = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;

    for (; t < n_tiles; t += nwarps) {
        const int4 d0 = __ldg(reinterpret_cast<const int4*>(tiles + t));
        const uint4 fl = __ldg(reinterpret_cast<const uint4*>(tiles + t) + 1);
        const long long base = ((long long)(unsigned)d0.x) | ((long long)d0.y << 32);
        const int row0 = d0.z;
        const int end = d0.w & 0xff, nrows = (d0.w >> 16) & 0xff;
        const unsigned F[4] = {fl.x, fl.y, fl.z, fl.w};
        // the descriptor of the tile after next: pulled into L2 while this tile is processed (costs no registers)
        const long long tn = (t + 2 * nwarps < n_tiles) ? t + 2 * nwarps : t;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(tiles + tn));

        if (end == 0) {
            // ---- long read (more than 128 entries, always ambiguous): the whole warp takes it
            const long long len = ((long long)F[0]) | ((long long)F[1] << 32);
            const long long hi = base + len;
            if (LONG8 && len <= 256) {
                // up to 8 entries per lane stay in registers: one pass over memory, like a regular tile
                int c8[8];
                double n8[8];
                double sum = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long p = base + lane + 32 * i;        // reads past the end hit padding / the next tile
                    c8[i] = ld_stream(col + p);
                    n8[i] = ld_stream(q + p);
                }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    n8[i] = (base + lane + 32 * i < hi) ? n8[i] * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, c8[i]) : 0.0;
                    sum += n8[i];
                }
                sum = group_sum<32>(sum, 0xffffffffu);
                const double g = (MODE == TILE_FUSED) ? a.wy[row0] * recip0(sum) : recip0(sum);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const long long p = base + lane + 32 * i;
                    if (p < hi) {
                        const double c = n8[i] * g;
                        if (MODE == TILE_FUSED) atomicAdd(my + c8[i], c);
                        if (MODE == TILE_Z) __stcs(a.z_out + p, c);
                        if (MODE == TILE_LNL) { if (c != 0.0) lnl_local += c * log1p_big(__ldg(q + p) * __ldg(a.inner_amb + c8[i]), s_log); }
                    }
                }
                continue;
            }
            // longer still: two passes over the read
            double sum = 0;
            for (long long p = base + lane; p < hi; p += 32)
                sum += ld_stream(q + p) * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, ld_stream(col + p));
            sum = group_sum<32>(sum, 0xffffffffu);
            const double g = (MODE == TILE_FUSED) ? a.wy[row0] * recip0(sum) : recip0(sum);
            if (MODE != TILE_FUSED || g != 0.0) {
                for (long long p = base + lane; p < hi; p += 32) {
                    const int cc = col[p];
                    const double qv = q[p];
                    const double c = (qv * gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc)) * g;
                    if (MODE == TILE_FUSED) { if (c != 0.0) atomicAdd(my + cc, c); }
                    if (MODE == TILE_Z) __stcs(a.z_out + p, c);
                    if (MODE == TILE_LNL) { if (c != 0.0) lnl_local += c * log1p_big(qv * __ldg(a.inner_amb + cc), s_log); }
                }
            }
            continue;
        }

        // ---- regular tile.  Lane-consecutive layout: lane l owns entries 32e + l
        const int* colp = col + base + lane;
        const double* qp = q + base + lane;
        int cc[4];
        double qq[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {      // 8 loads in flight before anything is consumed (entries past the tile
            cc[e] = ld_stream(colp + 32 * e);                    // are padding or the next tile's: harmless)
            qq[e] = ld_stream(qp + 32 * e);
        }
        // ... and the NEXT tile's entries go to L2 now (one lane asks the TMA unit: two bulk prefetches), so that the
        // loads above are L2 hits one iteration from now; its descriptor was prefetched an iteration ago
        if (TILE_PREFETCH && lane == 0 && t + nwarps < n_tiles) {
            const long long nb = __ldg(reinterpret_cast<const long long*>(tiles + t + nwarps));
            tile_prefetch_l2(q + nb, 128 * sizeof(double));
            tile_prefetch_l2(col + nb, 128 * sizeof(int));
        }
        // w*Y of my read (row phase below), requested early
        double w_mine = 1.0;
        if (MODE == TILE_FUSED) w_mine = (lane < nrows) ? __ldg(a.wy + row0 + lane) : 0.0;
        double n[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            double tv;
            if (MODE == TILE_FUSED) {
                tv = gather_pt<SMEM_TAB>(pt, s_tab, s_cols, cc[e]);
            } else {
                // a read with a single entry: a start here and a start right after (position 128 counts as one)
                const unsigned nxt = (e < 3) ? F[(e + 1) & 3] : 1u;
                const bool uniq = ((F[e] & ((F[e] >> 1) | (nxt << 31))) >> lane) & 1u;
                tv = __ldg((uniq ? a.tab_uni : pt) + cc[e]);
            }
            n[e] = ((32 * e + lane) < end) ? qq[e] * tv : 0.0;
            s_n[wl + 16 * e] = n[e];
        }
        __syncwarp();
        // Blocked layout for the row sums: lane l sums entries 4l .. 4l+3, one segmented scan over the 32 lane tails
        const double2 ma = *reinterpret_cast<const double2*>(s_n + 2 * lane);
        const double2 mb = *reinterpret_cast<const double2*>(s_n + 72 + 2 * lane);
        const double m[4] = {ma.x, ma.y, mb.x, mb.y};
        const unsigned w_lo = F[wsel];
        const unsigned w_hi = (wsel < 3) ? F[(wsel + 1) & 3] : 1u;          // position 128 counts as a read start
        const unsigned five = __funnelshift_r(w_lo, w_hi, sh) & 0x1fu;     // my 4 entries + the one after
        const int pc0 = __popc(F[0]), pc1 = pc0 + __popc(F[1]), pc2 = pc1 + __popc(F[2]);
        const int below = (wsel == 0 ? 0 : wsel == 1 ? pc0 : wsel == 2 ? pc1 : pc2) + __popc(w_lo & ((1u << sh) - 1u));
        double tail = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) { if ((five >> e) & 1u) tail = 0; tail += m[e]; }
        const unsigned heads = __ballot_sync(0xffffffffu, (five & 0xfu) != 0u);
        const unsigned hle = heads & le_mask;
        const int headlane = hle ? (31 - __clz(hle)) : 0;
        double x = tail;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double y = __shfl_up_sync(0xffffffffu, x, d);
            add_if_ge(x, y, lane - d, headlane);
        }
        double run = __shfl_up_sync(0xffffffffu, x, 1);
        if (lane == 0) run = 0.0;
        // an entry followed by a read start (or the tile end) closes its read: publish that read's total
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if ((five >> e) & 1u) run = 0.0;
            run += m[e];
            if (((five >> (e + 1)) & 1u) && (4 * lane + e) < end) s_g[below + __popc(five & ((2u << e) - 1u)) - 1] = run;
        }
        __syncwarp();
        // one lane per read: the per-read scale (FUSED: (w*Y) * recip0(total); otherwise recip0(total))
        if (lane < nrows) s_g[lane] = (w_mine != 0.0) ? w_mine * recip0(s_g[lane]) : 0.0;
        for (int rho = lane + 32; rho < nrows; rho += 32) {
            const double w = (MODE == TILE_FUSED) ? a.wy[row0 + rho] : 1.0;
            s_g[rho] = (w != 0.0) ? w * recip0(s_g[rho]) : 0.0;
        }
        __syncwarp();
        // back in the lane-consecutive layout: neighbouring lanes hit neighbouring loci, the REDs coalesce per sector
        int nb = 0;                        // read starts before the current round
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int lr = nb + __popc(F[e] & le_mask) - 1;
            nb += __popc(F[e]);
            if ((32 * e + lane) < end) {
                const double c = n[e] * s_g[lr];
                if (MODE == TILE_FUSED) atomicAdd(my + cc[e], c);     // adding an exact 0.0 (unique reads) is harmless
                if (MODE == TILE_Z) __stcs(a.z_out + base + 32 * e + lane, c);      // written once, never re-read here
                if (MODE == TILE_LNL) {
                    if (c != 0.0) {
                        const unsigned nxt = (e < 3) ? F[(e + 1) & 3] : 1u;
                        const bool uniq = ((F[e] & ((F[e] >> 1) | (nxt << 31))) >> lane) & 1u;
                        lnl_local += c * log1p_big(qq[e] * __ldg((uniq ? a.inner_uni : a.inner_amb) + cc[e]), s_log);
                    }
                }
            }
        }
        __syncwarp();   // scratch is reused by the next tile
    }
    if (MODE == TILE_LNL) {
        lnl_local = block_sum(lnl_local, s_red);
        if (threadIdx.x == 0) a.partials[blockIdx.x] = lnl_local;
    }
}

}  // namespace tsc
