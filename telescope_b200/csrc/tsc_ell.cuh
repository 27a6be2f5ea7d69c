// Fused E-step + M-step accumulation over a locus-clustered, sliced-ELL copy of the ambiguous reads -- the
// per-iteration hot kernel (round 2).  Replaces model.py:718-722 (E-step) + model.py:730-733 (M-step sums).
//
// Why a second layout.  The round-1 flat-tile kernel (tsc_tiles.cuh) is bound by the SM's L1 data pipe, not by HBM:
// per 120-entry tile it spends ~63 RED.ADD.F64 sector requests, ~36 gather lines and ~46 shared-memory wavefronts of
// the transpose/segmented scan (profiles/r1_fused_kernel_ncu_summary.md).  All three disappear when
//   * reads are CLUSTERED: ambiguous reads are sorted by (first locus, length) once at construction, so the reads a
//     warp sees back to back touch the same few dozen loci (a B200-sized matrix has ~30 000 entries per locus);
//   * a slice is 16 reads laid out ELL-style: lane (r, h) = (lane & 15, lane >> 4) owns entries 2t + h of read r,
//     t = 0..T-1.  A read's row sum is a private serial sum plus ONE shuffle -- no segmented scan, no transpose;
//   * each warp keeps a window of 128 loci in shared memory: pi*theta (gathers become conflict-free LDS) and SIXTEEN
//     private accumulator copies, one per read slot (row = locus, column = read slot, so the 16 lanes of a half-warp
//     always hit 16 different banks).  Loci are unique within a read, so the adds are plain LDS/DADD/STS -- no
//     atomics at all in the loop.  A window block is flushed (16 copies summed in fixed order, one coalesced
//     RED.ADD.F64 of 32 doubles) only when the sorted stream has moved past it: a few thousand sector REDs per
//     iteration instead of 5e8.
// The slice records form one byte stream in HBM; every warp takes one contiguous run of it (batches of 32 records: one
// index load per lane) -- byte-balanced at construction, re-partitioned by measured CTA time after the first iterations
// (k_ell_rebalance below: the byte-balanced runs end 10 % apart).  It asks the TMA
// unit to pull each record into L2 several records ahead (cp.async.bulk.prefetch.L2, one instruction per record) and
// then loads the record's loci / Q straight into registers -- every load of a slice is in flight at once, all of them
// L2 hits.  (A first version staged the records in a shared-memory ring with cp.async.bulk + mbarrier; ncu showed the
// shared-memory pipe ~75 % busy because the Q stream crossed it twice, TMA write + LDS, and the 12 KB ring capped the
// SM at 7 warps; see profiles/r2_ell_kernel_ncu_summary.md.)  With 17.6 KB of window per one-warp CTA, 12 warps per
// SM hide the L2 latency.
//
// Record (16-byte aligned, 128 + 288*T bytes):  double wy[16] | uint8 wcol[T][32] | double q[T][32]
// Index (16 bytes per record): byte offset / 16, first locus lo, T | last locus << 8.
// wcol = locus & 127, the entry's row in the window (blocks of 32 loci live in slot (locus >> 5) & 3), or 128 + (lane >> 4)
// for an empty slot: a dummy row per half-warp (every lane has its own dummy word), with q = 0 and pi*theta = 0, so
// empty slots add an exact +0.0 to a word nobody reads.
// With one byte per locus the stream moves ~9.5 B per entry where canonical CSR moves 12.
//
// Long reads (more than 2*kEllTMax = 48 entries; real multi-mappers of large TE families, the Zipf configuration) follow
// the slices in the same stream as ONE RECORD PER READ, `wy, pad | uint16 window-row[nch][32] | fp64 Q[nch][32]`, and
// have their own kernel (k_ell_long): the whole warp works on one read (lane l owns entries l, l+32, ...; the row sum is a
// warp reduction), the window is a single copy of 1024 loci in the same 17 KB of shared memory -- within one read the
// loci are distinct, so plain read-modify-write again, no atomics.
//
// Reads that fit neither form (a locus span above kEllSpan / kLongSpan, non-increasing loci) are not in the stream, and
// neither are the unique reads: those add nothing to the M-step sums (model.py:730-733; they enter pi through pisum0,
// model.py:699,738).  Both live in a compact residual CSR for the flat-tile kernels (ambiguous ones in front).
#pragma once
#include "tsc_kernels.cuh"

namespace tsc {

constexpr int kEllReads = 16;                 // reads per slice
constexpr int kEllTMax = 24;                  // steps per slice; a read holds at most 2*kEllTMax entries
constexpr int kEllWin = 128;                  // loci in a warp's window (4 blocks of 32)
constexpr int kEllSpan = 96;                  // max (locus - slice lo) inside a slice; lo % 32 + span < kEllWin
constexpr int kEllHdr = kEllReads * 8;        // wy
#ifndef TSC_ELL_AHEAD
#define TSC_ELL_AHEAD 6
#endif
constexpr int kEllAhead = TSC_ELL_AHEAD;      // records between the L2 prefetch and the loads
constexpr int kEllLenBits = 6;                // sort key = first locus << 6 | snake(length)
// long reads (more than 2*kEllTMax entries): one read per record, the whole warp on it, a single-copy window
constexpr int kLongWin = 512;                 // loci in the warp's window of the long-read kernel (16 blocks of 32): 8.5 KB per
                                              // one-warp CTA, so ~20 warps per SM hide the latency of one-read-at-a-time work
constexpr int kLongSpan = kLongWin - 32;      // max (last - first locus) of a long read
constexpr int kLongChunks = 127;              // chunks of 32 entries per long record (7 bits of the index word)
constexpr int kLongRegs = 8;                  // chunks kept in registers (reads of up to 256 entries: one pass)
static_assert(2 * kEllTMax < (1 << kEllLenBits), "length must fit the key");
static_assert(kEllTMax == 24, "k_ell_fused dispatches bodies of 4..24 steps");
static_assert(kEllAhead >= 1 && kEllAhead < 32, "prefetch distance is within one segment");

__host__ __device__ inline int ell_record_bytes(int T) { return kEllHdr + 288 * T; }
__host__ __device__ inline int ell_long_bytes(int nch) { return 16 + 320 * nch; }      // wy, pad | u16 row[nch][32] | q[nch][32]
// the T byte of an index entry: 1..kEllTMax = slice of 16 reads with T steps; 0x80 | nch = one long read of nch chunks
__host__ __device__ inline int ell_bytes_of(int tbyte) { return (tbyte & 0x80) ? ell_long_bytes(tbyte & 0x7f) : ell_record_bytes(tbyte); }

// ---------------------------------------------------------------------------------------------------------------
// exclusive prefix sums (construction only): out[i] = sum in[0..i), out[n] = total.  4096 items per block.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kScanItems = 4096;

template <typename T>
__global__ void __launch_bounds__(1024) k_scan_local(const T* __restrict__ in, long long n, long long* __restrict__ out,
                                                     long long* __restrict__ block_tot) {
    __shared__ long long s_w[32];
    const long long base = (long long)blockIdx.x * kScanItems + threadIdx.x * 4;
    long long v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = (base + i < n) ? (long long)in[base + i] : 0;
    const long long mine = v[0] + v[1] + v[2] + v[3];
    long long x = mine;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
    if (lane == 31) s_w[w] = x;
    __syncthreads();
    if (w == 0) {
        long long t = s_w[lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
        s_w[lane] = t;
    }
    __syncthreads();
    long long run = x - mine + (w ? s_w[w - 1] : 0);
#pragma unroll
    for (int i = 0; i < 4; ++i) { if (base + i < n) out[base + i] = run; run += v[i]; }
    if (threadIdx.x == 1023) block_tot[blockIdx.x] = s_w[31];
}

__global__ void __launch_bounds__(1024) k_scan_totals(long long* __restrict__ block_tot, int nb) {
    __shared__ long long s_w[32];
    __shared__ long long s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + threadIdx.x;
        const long long mine = (i < nb) ? block_tot[i] : 0;
        long long x = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_w[w] = x;
        __syncthreads();
        if (w == 0) {
            long long t = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
            s_w[lane] = t;
        }
        __syncthreads();
        const long long carry = s_carry;
        if (i < nb) block_tot[i] = carry + x - mine + (w ? s_w[w - 1] : 0);
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + s_w[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[nb] = s_carry;
}

__global__ void k_scan_add(long long* __restrict__ out, long long n, const long long* __restrict__ block_tot, int nb) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n; i += stride) out[i] += block_tot[i / kScanItems];
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = block_tot[nb];
}

// ---------------------------------------------------------------------------------------------------------------
// construction of the clustered stream
// ---------------------------------------------------------------------------------------------------------------

// Per read: the sort key of a stream candidate, or -1; candidates are counted per key.  Short reads (2..48 entries):
// (first locus, length in snake order, so that neighbouring keys hold reads of similar length) -> 16 of them share a
// slice.  Long reads (49..4064 entries, loci increasing, span <= kLongSpan): n_short_keys + first locus -> a record each.
__global__ void k_ell_classify(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ col, int n_short_keys,
                               int n_keys, int* __restrict__ key, unsigned* __restrict__ hist) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const long long s = ip[r], e = ip[r + 1];
        const long long len = e - s;
        int k = -1;
        if (len >= 2 && len <= 32 * kLongChunks) {
            const int first = col[s];
            bool ok = true;
            int prev = first;
            for (long long p = s + 1; p < e; ++p) { const int c = col[p]; ok = ok && (c > prev); prev = c; }
            if (ok && len <= 2 * kEllTMax && prev - first <= kEllSpan) {
                const int lk = (first & 1) ? ((1 << kEllLenBits) - 1 - (int)len) : (int)len;
                k = (first << kEllLenBits) | lk;
            } else if (ok && len > 2 * kEllTMax && prev - first <= kLongSpan) {
                k = n_short_keys + first;
            }
            if (k >= n_keys) k = -1;          // cannot happen for first < n_cols; keeps the histogram in bounds
        }
        key[r] = k;
        if (k >= 0) atomicAdd(hist + k, 1u);
    }
}

// Counting-sort scatter: candidates land in key order (order inside a key is whatever the atomics give; all reads
// of a key have the same first locus and length, so only the summation order inside a slice depends on it).
__global__ void k_ell_scatter(const int* __restrict__ key, long long n_rows, const long long* __restrict__ bin_start,
                              unsigned* __restrict__ cursor, int* __restrict__ sorted, int n_short_keys, int n_stream_keys,
                              long long long_pad) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const int k = key[r];
        if (k < 0 || k >= n_stream_keys) continue;
        long long pos = bin_start[k] + atomicAdd(cursor + k, 1u);
        if (k >= n_short_keys) pos += long_pad;        // the short slots are padded to whole slices
        sorted[pos] = (int)r;
    }
}

// Index entries of the long reads: one record each.
__global__ void k_ell_long_index(const long long* __restrict__ ip, const int* __restrict__ col, const int* __restrict__ sorted_long,
                                 long long n_long, int4* __restrict__ hdr, int* __restrict__ rec_bytes) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < n_long; i += stride) {
        const int r = sorted_long[i];
        const long long b = ip[r], e = ip[r + 1];
        const int nch = (int)((e - b + 31) >> 5);
        hdr[i] = make_int4(0, col[b], (0x80 | nch) | (col[e - 1] << 8), 1);
        rec_bytes[i] = ell_long_bytes(nch);
    }
}

// One warp per long read writes its record and completes its index entry.
__global__ void __launch_bounds__(256) k_ell_fill_long(const long long* __restrict__ ip, const int* __restrict__ col,
                                                       const double* __restrict__ q, const double* __restrict__ wy,
                                                       const int* __restrict__ sorted_long, long long n_long,
                                                       int4* __restrict__ hdr, const long long* __restrict__ rec_off,
                                                       unsigned char* __restrict__ stream) {
    const int lane = threadIdx.x & 31;
    long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; i < n_long; i += stride) {
        const int r = sorted_long[i];
        const long long b = ip[r], off = rec_off[i];
        const int len = (int)(ip[r + 1] - b), nch = (len + 31) >> 5;
        unsigned char* rec = stream + off;
        if (lane == 0) { hdr[i].x = (int)(off >> 4); reinterpret_cast<double*>(rec)[0] = wy[r]; reinterpret_cast<double*>(rec)[1] = 0.0; }
        unsigned short* wc = reinterpret_cast<unsigned short*>(rec + 16);
        double* qq = reinterpret_cast<double*>(rec + 16 + 64 * nch);
        for (int k = lane; k < nch * 32; k += 32) {
            const bool real = k < len;
            wc[k] = (unsigned short)(real ? (col[b + k] & (kLongWin - 1)) : (kLongWin + lane));      // empty: this lane's dummy word
            qq[k] = real ? q[b + k] : 0.0;
        }
    }
}

// One thread per slice of 16 consecutive sorted candidates: window base, steps, and eviction of the (rare) members
// whose last locus does not fit the window; evicted reads get key -2 and go to the residual CSR.
__global__ void k_ell_slices(const long long* __restrict__ ip, const int* __restrict__ col, int* __restrict__ sorted,
                             long long n_cand, long long n_slices, int* __restrict__ key, int4* __restrict__ hdr,
                             int* __restrict__ rec_bytes) {
    long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; s < n_slices; s += stride) {
        const long long p0 = s * kEllReads;
        const int lo = col[ip[sorted[p0]]];
        int hi = lo, T = 0, kept = 0;
        for (int m = 0; m < kEllReads; ++m) {
            const long long p = p0 + m;
            if (p >= n_cand) break;
            const int r = sorted[p];
            const long long b = ip[r], e = ip[r + 1];
            const int last = col[e - 1];
            if (last - lo > kEllSpan) { key[r] = -2; sorted[p] = -1; continue; }
            hi = max(hi, last);
            T = max(T, (int)((e - b + 1) >> 1));
            ++kept;
        }
        hdr[s] = make_int4(0, lo, T | (hi << 8), kept);
        rec_bytes[s] = ell_record_bytes(T);
    }
}

// Work split of the stream kernels: CTA w takes the records that START in [w, w+1) * total / n_ctas bytes -- one
// contiguous, byte-balanced run of the sorted stream per warp.  range[w] = first record of CTA w, range[n_ctas] = n.
__global__ void k_ell_ranges(const long long* __restrict__ rec_off, long long n_slices, int n_ctas, long long* __restrict__ range) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w > n_ctas) return;
    const long long base = rec_off[0], total = rec_off[n_slices] - base;
    const long long target = (w == n_ctas) ? total : (long long)((__int128)total * w / n_ctas);
    long long lo = 0, hi = n_slices;          // first record that starts at or after the target
    while (lo < hi) { const long long mid = (lo + hi) >> 1; if (rec_off[mid] - base < target) lo = mid + 1; else hi = mid; }
    range[w] = lo;
}

// One warp per slice writes its record and completes its index entry.  The 16 reads of a slice lie anywhere in the
// shard's CSR, so they are first copied read by read, lane-consecutive (coalesced), into shared memory and transposed
// into the record's (step, lane) layout from there: lane (r, h) taking q[b_r + 2t + h] straight from global memory read
// every 32-byte sector in two different steps and fetched 31 B of DRAM per entry instead of 12 (ncu, profiles/).
constexpr int kFillWarps = 4;
constexpr int kFillQStride = 2 * kEllTMax + 1;        // doubles per read: odd, so the 16 reads of a step hit 16 bank pairs
constexpr int kFillCStride = 2 * kEllTMax + 4;        // bytes per read: 13 words, injective modulo the 32 banks

__global__ void __launch_bounds__(kFillWarps * 32) k_ell_fill(const long long* __restrict__ ip, const int* __restrict__ col,
                                                              const double* __restrict__ q, const double* __restrict__ wy,
                                                              const int* __restrict__ sorted, long long n_cand, long long n_slices,
                                                              int4* __restrict__ hdr, const long long* __restrict__ rec_off,
                                                              unsigned char* __restrict__ stream) {
    __shared__ double s_q[kFillWarps][kEllReads * kFillQStride];
    __shared__ unsigned char s_c[kFillWarps][kEllReads * kFillCStride];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, r = lane & 15, h = lane >> 4;
    double* sq = s_q[wib];
    unsigned char* sc = s_c[wib];
    long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; s < n_slices; s += stride) {
        const int4 hd = hdr[s];
        const long long off = rec_off[s];
        unsigned char* rec = stream + off;
        __syncwarp();                                  // (the previous slice's shared-memory reads are done)
        if (lane == 0) hdr[s].x = (int)(off >> 4);
        const long long p = s * kEllReads + r;
        const int read = (p < n_cand) ? sorted[p] : -1;
        long long b = 0;
        int len = 0;
        if (read >= 0) { b = ip[read]; len = (int)(ip[read + 1] - b); }
        if (h == 0) reinterpret_cast<double*>(rec)[r] = (read >= 0) ? wy[read] : 0.0;
        // ---- read by read into shared memory (a slice member has at most 2 * kEllTMax = 48 entries)
#pragma unroll 4
        for (int m = 0; m < kEllReads; ++m) {
            const long long bm = __shfl_sync(0xffffffffu, b, m);
            const int lm = __shfl_sync(0xffffffffu, len, m);
            if (lane < lm) { sq[m * kFillQStride + lane] = q[bm + lane]; sc[m * kFillCStride + lane] = (unsigned char)(col[bm + lane] & (kEllWin - 1)); }
            if (lane + 32 < lm) {
                sq[m * kFillQStride + lane + 32] = q[bm + lane + 32];
                sc[m * kFillCStride + lane + 32] = (unsigned char)(col[bm + lane + 32] & (kEllWin - 1));
            }
        }
        __syncwarp();
        // ---- ... and out in the record's layout: lane (r, h) owns entries 2t + h of read r
        const int T = hd.z & 0xff;
        unsigned char* dc = rec + kEllHdr;
        double* qq = reinterpret_cast<double*>(rec + kEllHdr + 32 * T);
        for (int t = 0; t < T; ++t) {
            const int k = 2 * t + h;
            double qv = 0.0;
            int d = kEllWin + h;                                      // empty slot: this half-warp's dummy row
            if (k < len) { qv = sq[r * kFillQStride + k]; d = sc[r * kFillCStride + k]; }
            dc[t * 32 + lane] = (unsigned char)d;
            qq[t * 32 + lane] = qv;
        }
    }
}

// Residual CSR = the reads that are not in the stream: key < 0 or >= n_stream_keys (key == nullptr: everybody), unique
// reads included.
// counters[0] += ambiguous reads, [1] += their entries, [2] += (1 << kResShift | entries) per residual read,
// [3] += the same for the ambiguous residual reads.
constexpr int kResShift = 38;     // a cursor packs (reads << 38 | entries): < 2^26 reads and < 2^38 entries per GPU

__global__ void k_res_count(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ key, int n_stream_keys,
                            unsigned long long* __restrict__ counters) {
    unsigned long long rows = 0, ents = 0, res = 0, res_amb = 0;
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const long long len = ip[r + 1] - ip[r];
        if (len >= 2) { ++rows; ents += (unsigned long long)len; }
        if (key == nullptr || key[r] < 0 || key[r] >= n_stream_keys) {
            res += (1ULL << kResShift) | (unsigned long long)len;
            if (len >= 2) res_amb += (1ULL << kResShift) | (unsigned long long)len;
        }
    }
    // warp totals first: four atomics per warp
    for (int o = 16; o > 0; o >>= 1) {
        rows += __shfl_xor_sync(0xffffffffu, rows, o);
        ents += __shfl_xor_sync(0xffffffffu, ents, o);
        res += __shfl_xor_sync(0xffffffffu, res, o);
        res_amb += __shfl_xor_sync(0xffffffffu, res_amb, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (rows) atomicAdd(counters + 0, rows);
        if (ents) atomicAdd(counters + 1, ents);
        if (res) atomicAdd(counters + 2, res);
        if (res_amb) atomicAdd(counters + 3, res_amb);
    }
}

// One lane per read.  A residual read takes its place (read slot, first entry) from a packed atomic, so the read
// pointers come out increasing in slot order whatever order the reads arrive in.  Ambiguous reads fill the front of the
// residual (the fused kernel only visits those), unique reads the rest: cursor[1] starts where the front ends.  The
// unique reads -- a fifth of all reads, one entry each -- take their places with ONE atomic per warp (ballot + rank:
// ten million single atomics on the same word were most of this kernel's time) and copy their entry themselves; the
// ambiguous residual reads (few) are then copied by the whole warp, one read at a time.
__global__ void __launch_bounds__(256) k_res_append(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ col,
                                                    const double* __restrict__ q, const double* __restrict__ wy,
                                                    const int* __restrict__ key, int n_stream_keys,
                                                    unsigned long long* __restrict__ cursor,
                                                    long long* __restrict__ ip_out, int* __restrict__ col_out,
                                                    double* __restrict__ q_out, double* __restrict__ wy_out,
                                                    int* __restrict__ rowid_out) {
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long r_end = ((n_rows + stride - 1) / stride) * stride;       // every warp runs the same trip count
    for (; r < r_end; r += stride) {
        long long b = 0, e = 0;
        bool res = false;
        if (r < n_rows) { b = ip[r]; e = ip[r + 1]; res = (key == nullptr || key[r] < 0 || key[r] >= n_stream_keys); }
        const bool uni = res && (e - b) == 1, amb = res && (e - b) >= 2;
        // ---- unique reads: one packed atomic for the warp's reads, each lane places its own entry
        const unsigned um = __ballot_sync(0xffffffffu, uni);
        if (um) {
            const unsigned long long n = (unsigned long long)__popc(um);
            unsigned long long base = 0;
            if (lane == (__ffs(um) - 1)) base = atomicAdd(cursor + 1, (n << kResShift) | n);
            base = __shfl_sync(0xffffffffu, base, __ffs(um) - 1);
            if (uni) {
                const long long k = (long long)__popc(um & lt);
                const long long rp = (long long)(base >> kResShift) + k, o = (long long)(base & ((1ULL << kResShift) - 1ULL)) + k;
                ip_out[rp] = o; wy_out[rp] = wy[r]; rowid_out[rp] = (int)r;
                col_out[o] = col[b]; q_out[o] = q[b];
            }
        }
        // ---- ambiguous residual reads: the owner lane takes the place, the warp copies
        unsigned am = __ballot_sync(0xffffffffu, amb);
        unsigned long long old = 0;
        if (amb) old = atomicAdd(cursor + 0, (1ULL << kResShift) | (unsigned long long)(e - b));
        while (am) {
            const int src = __ffs(am) - 1;
            am &= am - 1;
            const long long bb = __shfl_sync(0xffffffffu, b, src), ee = __shfl_sync(0xffffffffu, e, src);
            const unsigned long long oo = __shfl_sync(0xffffffffu, old, src);
            const long long rr = __shfl_sync(0xffffffffu, r, src);
            const long long rp = (long long)(oo >> kResShift), o = (long long)(oo & ((1ULL << kResShift) - 1ULL));
            if (lane == 0) { ip_out[rp] = o; wy_out[rp] = wy[rr]; rowid_out[rp] = (int)rr; }
            for (long long k = bb + lane; k < ee; k += 32) { col_out[o + (k - bb)] = col[k]; q_out[o + (k - bb)] = q[k]; }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the stream kernels
// ---------------------------------------------------------------------------------------------------------------
// What a pass over the stream produces.
//   ELL_FUSED : E-step + M-step sums (the per-iteration kernel; model.py:718-722 + 730-733)
//   ELL_LNL   : log-likelihood of the stream's reads, sum z * log1p(Q * inner[locus]) with z from the E-step table
//               (model.py:744-760); the reads outside the stream go through k_tiles<TILE_LNL> on the residual CSR
//   ELL_REASSIGN : reassign(method, thresh, initial) column sums and best-hit counts of the stream's reads
//               (model.py:808-865, sparse_plus.py:99-165); lane = read makes row max / best-hit count / kept mass
//               private serial reductions plus one shuffle each
enum { ELL_FUSED = 0, ELL_LNL = 1, ELL_REASSIGN = 2 };

struct EllArgs {
    const unsigned char* stream;
    const int4* index;            // per record: byte offset / 16, lo, T | hi << 8, reads
    const long long* range;       // gridDim.x + 1 record boundaries: CTA w takes records [range[w], range[w+1])
    long long n_slices;
    const double* pt;             // pi*theta of the E-step (every read of the stream is ambiguous)
    double* acc;                  // FUSED: R replicas of K doubles
    int K, R;
    const EmState* st;            // nullptr = always run
    const double* inner;          // LNL: pi*theta inside log1p
    double* partials;             // LNL: one partial sum per CTA
    const LogTab* log_tab;        // LNL
    // REASSIGN: pt = the posterior's table (pi*theta of the last E-step, or all ones for Q.norm(1)); acc = colsum[K]
    int method;                   // tsc_method; TSC_CHOOSE counts the reads with a single best hit (like exclude)
    double thresh;
    const int* rowid;             // slices: 16 per record, the shard's read of each slot (-1 = empty); long reads: 1 per record
    int* nbest;                   // per read (optional): number of best hits
    unsigned* cta_ns;             // FUSED (optional): how long each CTA took, in ns -- input of k_ell_rebalance
};

struct EllReassign {              // REASSIGN: per-slice view handed to the body
    int method, want_colsum;
    double thresh;
    const int* rowid_slice;       // this slice's 16 reads
    int* nbest;
    int* s_cnt;                   // shared int window [kEllWin + 2] (exclude / all)
    double* s_avg;                // shared fp64 window [kEllWin + 2], single copy, CAS adds (average)
};

// Shared memory of a one-warp CTA of the slice kernels (doubles): acc [kEllWin + 2][16] | pt [kEllWin + 8], or
// (LNL) pt [kEllWin + 8] | inner [kEllWin + 8] | log table.  The long-read kernels: ell_long_smem_bytes.
constexpr int kShortDoubles = (kEllWin + 2) * kEllReads + kEllWin + 8;
constexpr int kLongDoubles = 2 * (kLongWin + 32);
template <int MODE>
__host__ __device__ constexpr size_t ell_smem_bytes() {
    return MODE != ELL_LNL ? sizeof(double) * kShortDoubles
                           : sizeof(double) * 2 * (kEllWin + 8) + sizeof(LogTab) * kLogTab;
}
constexpr size_t kEllSmem = ell_smem_bytes<ELL_FUSED>();

__device__ __forceinline__ void ell_prefetch_l2(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ double ell_ld_stream(const double* p) {      // read once: do not keep it in L1
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// LNL, out of line and practically never taken: the terms of a slice whose log1p argument is below 2^53 (or not
// finite), recomputed from the record with the library log1p.  Kept away from the main body's registers.
__device__ __noinline__ double ell_lnl_slow(const unsigned char* __restrict__ rec, int T, int lane, const double* s_pt,
                                            const double* s_in, double rr) {
    const unsigned char* cp = rec + kEllHdr + lane;
    const double* qp = reinterpret_cast<const double*>(rec + kEllHdr + 32 * T) + lane;
    double out = 0.0;
    for (int t = 0; t < T; ++t) {
        const unsigned jw = cp[32 * t];
        const double qv = qp[32 * t];
        const double z = (qv * s_pt[jw]) * rr, xt = qv * s_in[jw];
        if (z != 0.0 && !log_big_ok(xt)) out += z * log1p(xt);
    }
    return out;
}

// One slice record, TM = T rounded up to a multiple of 4: straight-line code, every load of the slice is in flight at
// once; only the last three steps are conditional.
template <int MODE, int TM>
__device__ __forceinline__ void ell_body(const unsigned char* __restrict__ rec, int T, int lane, const double* s_pt,
                                         unsigned char* accb /* FUSED: s_acc + 8 * (lane & 15) */,
                                         const double* s_in, const LogTab* s_log, double& lnl_local, const EllReassign& re) {
    const unsigned char* cp = rec + kEllHdr + lane;
    const double* qp = reinterpret_cast<const double*>(rec + kEllHdr + 32 * T) + lane;
    double w_mine = 1.0;
    if (MODE == ELL_FUSED) w_mine = __ldg(reinterpret_cast<const double*>(rec) + (lane & 15));
    double n[TM];
    unsigned ao[TM];              // window row of the entry, then (FUSED) byte offset of its accumulator row
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        n[t] = 0.0;
        ao[t] = kEllWin + (lane >> 4);                // this half-warp's dummy row
        if (t < TM - 3 || t < T) {
            ao[t] = __ldg(cp + 32 * t);
            n[t] = ell_ld_stream(qp + 32 * t);
        }
    }
    // ---- pass 1: numerators n = Q * (pi*theta)[locus], private row sum
    double x[MODE == ELL_LNL ? TM : 1];               // LNL: the argument of log1p, Q * inner[locus]
    double sp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        if (MODE == ELL_LNL) x[t] = n[t] * s_in[ao[t]];
        n[t] *= s_pt[ao[t]];
        sp[t & 3] += n[t];
        if (MODE == ELL_FUSED) ao[t] *= kEllReads * 8;
    }
    double sum = (sp[0] + sp[1]) + (sp[2] + sp[3]);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    if (MODE == ELL_FUSED) {
        // (w*Y) * recip0(total), the same expression as the tile kernel; empty read slots have w = 0
        const double g = (w_mine != 0.0) ? w_mine * recip0(sum) : 0.0;
        // ---- pass 2: c = n * g into this read slot's private accumulator column.  Loci are unique within a read, the
        // two lanes of a read hold different entries, and empty slots point at the dummy row, so no two real updates of
        // a slice share an address: all loads may precede all stores.
#pragma unroll
        for (int t = 0; t < TM; ++t) n[t] = *reinterpret_cast<const double*>(accb + ao[t]) + n[t] * g;
#pragma unroll
        for (int t = 0; t < TM; ++t) *reinterpret_cast<double*>(accb + ao[t]) = n[t];
        __syncwarp();     // the other lane of this read slot loads these words in the next slice
    } else if (MODE == ELL_REASSIGN) {
        // z of the read, its maximum, how many entries reach it, and (conf) the mass at or above the threshold: private
        // over the lane's entries, then one exchange with the other lane of the read
        const double rr = recip0(sum);
        double zmax = 0.0, kept = 0.0;
#pragma unroll
        for (int t = 0; t < TM; ++t) {
            n[t] *= rr;
            zmax = fmax(zmax, n[t]);
            if (n[t] >= re.thresh) kept += n[t];
        }
        zmax = fmax(zmax, __shfl_xor_sync(0xffffffffu, zmax, 16));
        kept += __shfl_xor_sync(0xffffffffu, kept, 16);
        int nb = 0;
#pragma unroll
        for (int t = 0; t < TM; ++t) nb += (n[t] == zmax && n[t] != 0.0);
        nb += __shfl_xor_sync(0xffffffffu, nb, 16);
        if (re.nbest && lane < 16 && re.rowid_slice[lane] >= 0) re.nbest[re.rowid_slice[lane]] = nb;
        if (re.want_colsum) {
            if (re.method == 3) {                        // conf: every surviving entry, renormalised -> dense private columns
                const double rk = recip0(kept);
#pragma unroll
                for (int t = 0; t < TM; ++t) {
                    const double v = (n[t] >= re.thresh) ? n[t] * rk : 0.0;
                    n[t] = *reinterpret_cast<const double*>(accb + ao[t] * (kEllReads * 8)) + v;
                }
#pragma unroll
                for (int t = 0; t < TM; ++t) *reinterpret_cast<double*>(accb + ao[t] * (kEllReads * 8)) = n[t];
                __syncwarp();
            } else if (re.method == 2) {                 // average: 1/nbest on every best hit (a few per read)
                const double v = nb > 0 ? 1.0 / (double)nb : 0.0;
#pragma unroll
                for (int t = 0; t < TM; ++t)
                    if (n[t] == zmax && n[t] != 0.0) atomicAdd(re.s_avg + ao[t], v);
            } else if (re.method == 5) {                 // all: one per stored entry with a positive posterior
#pragma unroll
                for (int t = 0; t < TM; ++t)
                    if (n[t] > 0.0) atomicAdd(re.s_cnt + ao[t], 1);
            } else if (re.method <= 1) {                 // exclude (and the untied part of choose): the single best hit
#pragma unroll
                for (int t = 0; t < TM; ++t)
                    if (nb == 1 && n[t] == zmax && n[t] != 0.0) atomicAdd(re.s_cnt + ao[t], 1);
            }                                            // unique (4): every read of the stream is ambiguous -> nothing
        }
    } else {
        // straight-line: every entry takes the table-driven log; entries whose argument is outside its range (never,
        // in practice) are redone with the library log1p afterwards
        const double rr = recip0(sum);
        double acc[2] = {0.0, 0.0};
        bool slow = false;
#pragma unroll
        for (int t = 0; t < TM; ++t) {
            const double xt = x[MODE == ELL_LNL ? t : 0];
            n[t] *= rr;                                           // z
            const double term = n[t] * log_big_core(xt, s_log);
            const bool use = n[t] != 0.0, ok = log_big_ok(xt);
            acc[t & 1] += (use && ok) ? term : 0.0;
            slow = slow || (use && !ok);
        }
        lnl_local += acc[0] + acc[1];
        if (__any_sync(0xffffffffu, slow)) lnl_local += ell_lnl_slow(rec, T, lane, s_pt, s_in, rr);
    }
}

// ---- long reads: one read per record, lane l owns entries l, l + 32, ...; window = single-copy arrays of kLongWin loci
struct EllLongWin {
    double* acc;                  // [kLongWin + 32] (FUSED / REASSIGN); the last 32 = one dummy word per lane
    const double* pt;             // [kLongWin + 32], dummies = 0
    const double* in;             // LNL: [kLongWin + 32]
};

__device__ __forceinline__ double warp_sum(double v) { return group_sum<32>(v, 0xffffffffu); }
__device__ __forceinline__ double warp_max(double v) { return group_max<32>(v, 0xffffffffu); }
__device__ __forceinline__ int warp_sum_int(int v) { return group_sum_int<32>(v, 0xffffffffu); }

// what a long read adds at an entry with posterior z (REASSIGN; model.py:837-862)
__device__ __forceinline__ double ell_reassign_value(int method, double z, double zmax, int nb, double thresh, double rkept) {
    const bool best = (z == zmax && z != 0.0);
    switch (method) {
        case 0: case 1: return (best && nb == 1) ? 1.0 : 0.0;
        case 2: return best ? 1.0 / (double)nb : 0.0;
        case 3: return (z >= thresh) ? z * rkept : 0.0;
        case 5: return (z > 0.0) ? 1.0 : 0.0;
        default: return 0.0;          // unique: a long read is ambiguous
    }
}

// nch <= NC chunks, everything in registers: one pass over the record, in two phases so that two reads can be in
// flight per warp (their load and reduction chains are independent; only the window updates are ordered).
template <int MODE, int NC>
struct EllLongState {
    double n[NC];
    unsigned row[NC];
    double x[MODE == ELL_LNL ? NC : 1];
    double sum;
};

// phase 1: loads, numerators n = Q * (pi*theta)[locus], this lane's part of the row sum
template <int MODE, int NC>
__device__ __forceinline__ void ell_long_p1(EllLongState<MODE, NC>& st, const unsigned char* __restrict__ rec, int nch, int lane,
                                            const EllLongWin w) {
    const unsigned short* wp = reinterpret_cast<const unsigned short*>(rec + 16) + lane;
    const double* qp = reinterpret_cast<const double*>(rec + 16 + 64 * nch) + lane;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        st.n[c] = 0.0;
        st.row[c] = kLongWin + lane;                    // this lane's dummy word
        if (c < 2 || c < nch) { st.row[c] = __ldg(wp + 32 * c); st.n[c] = ell_ld_stream(qp + 32 * c); }
    }
    st.sum = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (MODE == ELL_LNL) st.x[c] = st.n[c] * w.in[st.row[c]];
        st.n[c] *= w.pt[st.row[c]];
        st.sum += st.n[c];
    }
}

// phase 2 (st.sum already reduced over the warp): the read's contribution
template <int MODE, int NC>
__device__ __forceinline__ void ell_long_p2(EllLongState<MODE, NC>& st, const unsigned char* __restrict__ rec, int nch, int lane,
                                            const EllLongWin w, const LogTab* s_log, double& lnl_local, const EllReassign& re,
                                            const int* rowid) {
    if (MODE == ELL_FUSED) {
        const double wy = __ldg(reinterpret_cast<const double*>(rec));
        const double g = (wy != 0.0) ? wy * recip0(st.sum) : 0.0;
        // loci are unique within a read and every empty slot has its own dummy word: plain read-modify-write
#pragma unroll
        for (int c = 0; c < NC; ++c) st.n[c] = w.acc[st.row[c]] + st.n[c] * g;
#pragma unroll
        for (int c = 0; c < NC; ++c) w.acc[st.row[c]] = st.n[c];
    } else if (MODE == ELL_REASSIGN) {
        const double rr = recip0(st.sum);
        double zmax = 0.0, kept = 0.0;
#pragma unroll
        for (int c = 0; c < NC; ++c) { st.n[c] *= rr; zmax = fmax(zmax, st.n[c]); if (st.n[c] >= re.thresh) kept += st.n[c]; }
        zmax = warp_max(zmax);
        kept = warp_sum(kept);
        int nb = 0;
#pragma unroll
        for (int c = 0; c < NC; ++c) nb += (st.n[c] == zmax && st.n[c] != 0.0);
        nb = warp_sum_int(nb);
        if (re.nbest && lane == 0 && rowid[0] >= 0) re.nbest[rowid[0]] = nb;
        if (re.want_colsum) {
            const double rk = recip0(kept);
#pragma unroll
            for (int c = 0; c < NC; ++c) st.n[c] = w.acc[st.row[c]] + ell_reassign_value(re.method, st.n[c], zmax, nb, re.thresh, rk);
#pragma unroll
            for (int c = 0; c < NC; ++c) w.acc[st.row[c]] = st.n[c];
        }
    } else {
        const double rr = recip0(st.sum);
        double acc2[2] = {0.0, 0.0};
        bool slow = false;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double xt = st.x[MODE == ELL_LNL ? c : 0];
            st.n[c] *= rr;
            const double term = st.n[c] * log_big_core(xt, s_log);
            const bool use = st.n[c] != 0.0, ok = log_big_ok(xt);
            acc2[c & 1] += (use && ok) ? term : 0.0;
            slow = slow || (use && !ok);
        }
        lnl_local += acc2[0] + acc2[1];
        if (__any_sync(0xffffffffu, slow)) {
            const unsigned short* wp = reinterpret_cast<const unsigned short*>(rec + 16) + lane;
            const double* qp = reinterpret_cast<const double*>(rec + 16 + 64 * nch) + lane;
#pragma unroll 1
            for (int c = 0; c < nch; ++c) {             // rare: recomputed from the record, library log1p
                const unsigned r2 = wp[32 * c];
                const double qv = qp[32 * c], z = (qv * w.pt[r2]) * rr, xt = qv * w.in[r2];
                if (z != 0.0 && !log_big_ok(xt)) lnl_local += z * log1p(xt);
            }
        }
    }
}

template <int MODE, int NC>
__device__ __forceinline__ void ell_long(const unsigned char* __restrict__ rec, int nch, int lane, const EllLongWin w,
                                         const LogTab* s_log, double& lnl_local, const EllReassign re, const int* rowid) {
    EllLongState<MODE, NC> st;
    ell_long_p1<MODE, NC>(st, rec, nch, lane, w);
    st.sum = warp_sum(st.sum);
    ell_long_p2<MODE, NC>(st, rec, nch, lane, w, s_log, lnl_local, re, rowid);
}

// two reads at once (both at most NC chunks, both inside the window)
template <int MODE, int NC>
__device__ __forceinline__ void ell_long2(const unsigned char* __restrict__ recA, int nchA, const unsigned char* __restrict__ recB,
                                          int nchB, int lane, const EllLongWin w, const LogTab* s_log, double& lnl_local,
                                          const EllReassign re, const int* rowidA) {
    EllLongState<MODE, NC> A, B;
    ell_long_p1<MODE, NC>(A, recA, nchA, lane, w);
    ell_long_p1<MODE, NC>(B, recB, nchB, lane, w);
    A.sum = warp_sum(A.sum);
    B.sum = warp_sum(B.sum);
    ell_long_p2<MODE, NC>(A, recA, nchA, lane, w, s_log, lnl_local, re, rowidA);
    if (MODE != ELL_LNL) __syncwarp();               // B may touch the window words A just wrote, from other lanes
    ell_long_p2<MODE, NC>(B, recB, nchB, lane, w, s_log, lnl_local, re, rowidA + 1);
}

// more than kLongRegs chunks (reads above 256 entries): the record is walked once per reduction (L2 hits)
template <int MODE>
__device__ __noinline__ void ell_long_big(const unsigned char* __restrict__ rec, int nch, int lane, const EllLongWin w,
                                          const LogTab* s_log, double& lnl_local, const EllReassign re, const int* rowid) {
    const unsigned short* wp = reinterpret_cast<const unsigned short*>(rec + 16) + lane;
    const double* qp = reinterpret_cast<const double*>(rec + 16 + 64 * nch) + lane;
    double sum = 0.0;
    for (int c = 0; c < nch; ++c) sum += qp[32 * c] * w.pt[wp[32 * c]];
    sum = warp_sum(sum);
    const double rr = recip0(sum);
    if (MODE == ELL_FUSED) {
        const double wy = __ldg(reinterpret_cast<const double*>(rec));
        const double g = (wy != 0.0) ? wy * rr : 0.0;
        for (int c = 0; c < nch; ++c) { const unsigned r2 = wp[32 * c]; w.acc[r2] += (qp[32 * c] * w.pt[r2]) * g; }
    } else if (MODE == ELL_LNL) {
        for (int c = 0; c < nch; ++c) {
            const unsigned r2 = wp[32 * c];
            const double qv = qp[32 * c], z = (qv * w.pt[r2]) * rr;
            if (z != 0.0) lnl_local += z * log1p_big(qv * w.in[r2], s_log);
        }
    } else {
        double zmax = 0.0, kept = 0.0;
        for (int c = 0; c < nch; ++c) {
            const double z = (qp[32 * c] * w.pt[wp[32 * c]]) * rr;
            zmax = fmax(zmax, z);
            if (z >= re.thresh) kept += z;
        }
        zmax = warp_max(zmax);
        kept = warp_sum(kept);
        int nb = 0;
        for (int c = 0; c < nch; ++c) { const double z = (qp[32 * c] * w.pt[wp[32 * c]]) * rr; nb += (z == zmax && z != 0.0); }
        nb = warp_sum_int(nb);
        if (re.nbest && lane == 0 && rowid[0] >= 0) re.nbest[rowid[0]] = nb;
        if (re.want_colsum) {
            const double rk = recip0(kept);
            for (int c = 0; c < nch; ++c) {
                const unsigned r2 = wp[32 * c];
                w.acc[r2] += ell_reassign_value(re.method, (qp[32 * c] * w.pt[r2]) * rr, zmax, nb, re.thresh, rk);
            }
        }
    }
}

__device__ __forceinline__ unsigned long long ell_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

#ifdef TSC_ELL_TRACE
// debug build only (tools/profile/trace_ell.py): per-CTA start / end times of the last k_ell<ELL_FUSED> launch
__device__ unsigned long long g_ell_trace[2 * 8192];
__device__ __forceinline__ unsigned long long ell_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#endif

template <int MODE>
__global__ void __launch_bounds__(32) k_ell(const EllArgs a) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    if (a.st && a.st->done) return;
#ifdef TSC_ELL_TRACE
    if (MODE == ELL_FUSED && threadIdx.x == 0 && blockIdx.x < 8192) g_ell_trace[2 * blockIdx.x] = ell_now();
#endif
    unsigned long long t_start = 0;
    if (MODE == ELL_FUSED && a.cta_ns) t_start = ell_globaltimer();
    // FUSED: s_acc [kEllWin + 2][kEllReads] (last two rows = dummies) | s_pt [kEllWin + 8] ([kEllWin..] = 0 for empty slots)
    // LNL:   s_pt [kEllWin + 8] | s_in [kEllWin + 8] | log table
    double* s_acc = reinterpret_cast<double*>(s_raw);
    double* s_pt = (MODE != ELL_LNL) ? s_acc + (kEllWin + 2) * kEllReads : s_acc;
    double* s_in = s_pt + kEllWin + 8;
    LogTab* s_log = reinterpret_cast<LogTab*>(s_in + kEllWin + 8);
    const int lane = threadIdx.x;
    const long long r_begin = a.range[blockIdx.x], r_end = a.range[blockIdx.x + 1];     // this warp's run of records
    const int K = a.K;
    const double* __restrict__ pt = a.pt;
    const unsigned char* __restrict__ stream = a.stream;
    double* my = (MODE != ELL_LNL && a.acc) ? a.acc + (size_t)(blockIdx.x % a.R) * K : nullptr;
    unsigned char* accb = reinterpret_cast<unsigned char*>(s_acc) + 8 * (lane & 15);
    double lnl_local = 0.0;
    // REASSIGN: the count / average windows overlay the accumulator columns (one method per launch)
    EllReassign re{a.method, a.acc != nullptr, a.thresh, nullptr, a.nbest, reinterpret_cast<int*>(s_acc), s_acc + kEllWin + 8};

    for (int i = lane; i < (int)(ell_smem_bytes<MODE>() / 8); i += 32) s_acc[i] = 0.0;     // accumulators, tables
    if (MODE == ELL_LNL) {
        __syncwarp();
        for (int i = lane; i < kLogTab; i += 32) s_log[i] = a.log_tab[i];
    }
    __syncwarp();

    // lane l of a batch's index registers describes record b + l (T = 0 beyond the end of the run)
    auto load_batch = [&](long long b) -> int4 {
        int4 v = make_int4(0, 0, 0, 0);
        if (b + lane < r_end) v = __ldg(a.index + b + lane);
        return v;
    };
    auto prefetch_mine = [&](const int4& v) {          // this lane's record -> L2
        const int T = v.z & 0xff;
        if (T) ell_prefetch_l2(stream + ((long long)(unsigned)v.x << 4), (unsigned)ell_record_bytes(T));
    };
    auto flush_block = [&](int b) {
        double s = 0.0;
        if (MODE == ELL_REASSIGN && a.method != 3) {         // single-copy windows: counts (exclude, all) or 1/nbest sums
            const int w = (b & 3) * 32 + lane;
            if (a.method == 2) { s = re.s_avg[w]; re.s_avg[w] = 0.0; }
            else { s = (double)re.s_cnt[w]; re.s_cnt[w] = 0; }
        } else {
            const int row = ((b & 3) * 32 + lane) * kEllReads;
#pragma unroll
            for (int c = 0; c < kEllReads; ++c) s += s_acc[row + ((c + lane) & 15)];
#pragma unroll
            for (int c = 0; c < kEllReads; ++c) s_acc[row + ((c + lane) & 15)] = 0.0;
        }
        const int j = b * 32 + lane;
        if (my && j < K && s != 0.0) atomicAdd(my + j, s);
    };

    int4 cur = load_batch(r_begin), nxt = load_batch(r_begin + 32);
    if (lane < kEllAhead) prefetch_mine(cur);         // the first records of the run
    int Fb = -1;                                      // first block (32 loci) of the window; -1 = empty

    for (long long b = r_begin; b < r_end; b += 32) {
        const int4 after = load_batch(b + 64);        // requested two batches early
        const int nrec = (int)min(32LL, r_end - b);
        for (int c = 0; c < nrec; ++c) {
            // ---- the lane that owns record c + kEllAhead (of this batch or the next) sends it to L2
            {
                const int pc = c + kEllAhead;
                if (lane == (pc & 31)) prefetch_mine(pc < 32 ? cur : nxt);
            }
            const unsigned off16 = (unsigned)__shfl_sync(0xffffffffu, cur.x, c);
            const int lo = __shfl_sync(0xffffffffu, cur.y, c);
            const int thi = __shfl_sync(0xffffffffu, cur.z, c);
            const int T = thi & 0xff, hi = thi >> 8;
            const unsigned char* rec = stream + ((long long)off16 << 4);

            // ---- window: blocks [Fb, Fb+4) of 32 loci; the stream is sorted by lo, so it only moves forward
            const int lb = lo >> 5;
            if (Fb < 0 || (hi >> 5) >= Fb + 4) {
                __syncwarp();
                int first_new = lb;
                if (Fb >= 0) {
                    if (MODE != ELL_LNL) {
                        const int e = min(lb, Fb + 4);
                        for (int blk = Fb; blk < e; ++blk) flush_block(blk);
                    }
                    first_new = max(lb, Fb + 4);
                }
                for (int blk = first_new; blk < lb + 4; ++blk) {
                    const int j = blk * 32 + lane;
                    s_pt[(blk & 3) * 32 + lane] = (j < K) ? __ldg(pt + j) : 0.0;
                    if (MODE == ELL_LNL) s_in[(blk & 3) * 32 + lane] = (j < K) ? __ldg(a.inner + j) : 0.0;
                }
                Fb = lb;
                __syncwarp();
            }
            if (MODE == ELL_REASSIGN) re.rowid_slice = a.rowid + (b + c) * kEllReads;
            switch ((T + 3) >> 2) {
                case 1: ell_body<MODE, 4>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
                case 2: ell_body<MODE, 8>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
                case 3: ell_body<MODE, 12>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
                case 4: ell_body<MODE, 16>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
                case 5: ell_body<MODE, 20>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
                default: ell_body<MODE, 24>(rec, T, lane, s_pt, accb, s_in, s_log, lnl_local, re); break;
            }
        }
        cur = nxt;
        nxt = after;
    }
    if (MODE != ELL_LNL) {
        // ---- run done: hand the window to the global accumulator
        __syncwarp();
        if (Fb >= 0) for (int blk = Fb; blk < Fb + 4; ++blk) flush_block(blk);
    }
    if (MODE == ELL_LNL) {
        lnl_local = group_sum<32>(lnl_local, 0xffffffffu);
        if (lane == 0) a.partials[blockIdx.x] = lnl_local;
    }
#ifdef TSC_ELL_TRACE
    if (MODE == ELL_FUSED && threadIdx.x == 0 && blockIdx.x < 8192) g_ell_trace[2 * blockIdx.x + 1] = ell_now();
#endif
    if (MODE == ELL_FUSED && a.cta_ns && threadIdx.x == 0) a.cta_ns[blockIdx.x] = (unsigned)(ell_globaltimer() - t_start);
}

// Measured re-partition of the slice stream.  The byte-balanced runs of k_ell_ranges do not finish together: how fast a
// warp gets through its run depends on the loci it covers (how often the window moves, how long the slices are), and the
// slowest warp of the 1776 sets the kernel time -- a CTA trace on B200 shows the last run ending 11 % after the median
// one, at every problem size (profiles/r2_cta_trace.md).  So the per-iteration kernel reports how long every CTA took and
// this kernel, run after the first iterations of a model, moves the boundaries to where equal shares of the MEASURED time
// fall: time is taken as uniform per byte inside an old run, new boundary k is the first record at or after the byte
// where k/G of the total time has passed.  One block; range[] is rewritten in place for the next launch.
__global__ void __launch_bounds__(1024) k_ell_rebalance(const int4* __restrict__ index, long long n_slices, unsigned end_off16,
                                                        int G, long long* __restrict__ range, const unsigned* __restrict__ cta_ns,
                                                        double* __restrict__ scratch /* 2 * (G + 1) */, const EmState* __restrict__ st) {
    if (st && st->done) return;                      // the launch that would have measured was skipped too
    double* S = scratch;                              // S[w] = time of the runs before w
    double* B = scratch + G + 1;                      // B[w] = first byte / 16 of run w
    for (int w = threadIdx.x; w <= G; w += blockDim.x) {
        const long long r = range[w];
        B[w] = (r >= n_slices) ? (double)end_off16 : (double)(unsigned)index[r].x;
    }
    {   // exclusive prefix sums of the run times: four consecutive runs per thread, warp scan, scan of the warp totals
        __shared__ double s_w[32];
        const int w0 = threadIdx.x * 4;
        double v[4], mine = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = (w0 + i < G) ? (double)max(cta_ns[w0 + i], 1u) : 0.0; mine += v[i]; }
        double x = mine;
        const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const double y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        if (lane == 31) s_w[wp] = x;
        __syncthreads();
        if (wp == 0) {
            double t = s_w[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { const double y = __shfl_up_sync(0xffffffffu, t, d); if (lane >= d) t += y; }
            s_w[lane] = t;
        }
        __syncthreads();
        double run = x - mine + (wp ? s_w[wp - 1] : 0.0);
#pragma unroll
        for (int i = 0; i < 4; ++i) { if (w0 + i < G) S[w0 + i] = run; run += v[i]; }
        if (threadIdx.x == 1023) S[G] = s_w[31];
    }
    __syncthreads();
    const double total = S[G];
    long long mine[4];                                // G <= 4096 boundaries, computed before anything is overwritten
    int nm = 0;
    for (int k = threadIdx.x; k <= G && nm < 4; k += blockDim.x, ++nm) {
        long long rec;
        if (k == 0) rec = 0;
        else if (k == G) rec = n_slices;
        else {
            const double tau = total * (double)k / (double)G;
            int lo = 0, hi = G - 1;                   // last run w with S[w] <= tau
            while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (S[mid] <= tau) lo = mid; else hi = mid - 1; }
            const double span = S[lo + 1] - S[lo];
            const double frac = span > 0.0 ? (tau - S[lo]) / span : 0.0;
            const double byte16 = B[lo] + frac * (B[lo + 1] - B[lo]);
            long long a = 0, b = n_slices;            // first record that starts at or after that byte
            while (a < b) { const long long mid = (a + b) >> 1; if ((double)(unsigned)index[mid].x < byte16) a = mid + 1; else b = mid; }
            rec = a;
        }
        mine[nm] = rec;
    }
    __syncthreads();
    nm = 0;
    for (int k = threadIdx.x; k <= G && nm < 4; k += blockDim.x, ++nm) range[k] = mine[nm];
}


// ---- the same passes over the long-read records (their own launch: the slice kernel stays free of the mode logic)
template <int MODE>
__host__ __device__ constexpr size_t ell_long_smem_bytes() {
    return sizeof(double) * kLongDoubles + (MODE == ELL_LNL ? sizeof(LogTab) * kLogTab : 0);
}

template <int MODE>
__global__ void __launch_bounds__(32) k_ell_long(const EllArgs a) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    if (a.st && a.st->done) return;
    // FUSED / REASSIGN: acc [kLongWin + 32] | pt [kLongWin + 32];   LNL: pt | inner | log table
    double* l_acc = reinterpret_cast<double*>(s_raw);
    double* l_pt = (MODE != ELL_LNL) ? l_acc + kLongWin + 32 : l_acc;
    double* l_in = l_pt + kLongWin + 32;
    LogTab* s_log = reinterpret_cast<LogTab*>(l_acc + kLongDoubles);
    const int lane = threadIdx.x;
    // Work split: batches of 32 consecutive records, dealt round-robin to the warps -- at any moment the resident warps
    // read neighbouring batches, i.e. a few pages of the stream (a contiguous run per warp would keep ~3000 distant pages
    // live at once; with ~1.5 KB per record that costs a TLB miss per read).  The stream is sorted by first locus, so a
    // warp's consecutive batches still move forward through the loci.
    const long long r_end = a.n_slices;                       // (here: the number of long-read records)
    const long long b_step = (long long)gridDim.x * 32;
    const int K = a.K;
    const double* __restrict__ pt = a.pt;
    const unsigned char* __restrict__ stream = a.stream;
    double* my = (MODE != ELL_LNL && a.acc) ? a.acc + (size_t)(blockIdx.x % a.R) * K : nullptr;
    double lnl_local = 0.0;
    EllReassign re{a.method, a.acc != nullptr, a.thresh, nullptr, a.nbest, nullptr, nullptr};
    const EllLongWin lw{l_acc, l_pt, l_in};
    constexpr int kBlocks = kLongWin / 32;

    for (int i = lane; i < kLongDoubles; i += 32) l_acc[i] = 0.0;
    if (MODE == ELL_LNL)
        for (int i = lane; i < kLogTab; i += 32) s_log[i] = a.log_tab[i];
    __syncwarp();

    auto load_batch = [&](long long b) -> int4 {
        int4 v = make_int4(0, 0, 0, 0);
        if (b + lane < r_end) v = __ldg(a.index + b + lane);
        return v;
    };
    auto prefetch_mine = [&](const int4& v) {
        const int tb = v.z & 0xff;
        if (tb) ell_prefetch_l2(stream + ((long long)(unsigned)v.x << 4), (unsigned)ell_bytes_of(tb));
    };
    auto flush_block = [&](int blk) {
        const int w = (blk & (kBlocks - 1)) * 32 + lane;
        const double s = l_acc[w];
        l_acc[w] = 0.0;
        const int j = blk * 32 + lane;
        if (my && j < K && s != 0.0) atomicAdd(my + j, s);
    };

    const long long b_first = (long long)blockIdx.x * 32;
    int4 cur = load_batch(b_first), nxt = load_batch(b_first + b_step);
    if (lane < kEllAhead) prefetch_mine(cur);
    int Fb = -1;                                      // first block (32 loci) of the window; -1 = empty

    for (long long b = b_first; b < r_end; b += b_step) {
        const int4 after = load_batch(b + 2 * b_step);
        const int nrec = (int)min(32LL, r_end - b);
        for (int c = 0; c < nrec; ++c) {
            {
                const int pc = c + kEllAhead;
                if (lane == (pc & 31)) prefetch_mine(pc < 32 ? cur : nxt);
            }
            const unsigned off16 = (unsigned)__shfl_sync(0xffffffffu, cur.x, c);
            const int lo = __shfl_sync(0xffffffffu, cur.y, c);
            const int thi = __shfl_sync(0xffffffffu, cur.z, c);
            const int nch = thi & 0x7f, hi = thi >> 8;
            const unsigned char* rec = stream + ((long long)off16 << 4);
            // ---- window: blocks [Fb, Fb + 32) of 32 loci; the long reads are sorted by first locus
            const int lb = lo >> 5;
            if (Fb < 0 || (hi >> 5) >= Fb + kBlocks) {
                __syncwarp();
                int first_new = lb;
                if (Fb >= 0) {
                    if (MODE != ELL_LNL) {
                        const int e = min(lb, Fb + kBlocks);
                        for (int blk = Fb; blk < e; ++blk) flush_block(blk);
                    }
                    first_new = max(lb, Fb + kBlocks);
                }
                for (int blk = first_new; blk < lb + kBlocks; ++blk) {
                    const int j = blk * 32 + lane, w = (blk & (kBlocks - 1)) * 32 + lane;
                    l_pt[w] = (j < K) ? __ldg(pt + j) : 0.0;
                    if (MODE == ELL_LNL) l_in[w] = (j < K) ? __ldg(a.inner + j) : 0.0;
                }
                Fb = lb;
                __syncwarp();
            }
            const int* rowid = (MODE == ELL_REASSIGN) ? a.rowid + (b + c) : nullptr;
            // ---- the next read of the batch goes along when it fits the registers and the window as it stands
            if (nch <= kLongRegs && c + 1 < nrec) {
                const int thi2 = __shfl_sync(0xffffffffu, cur.z, c + 1);
                const int nch2 = thi2 & 0x7f;
                if (nch2 <= kLongRegs && ((thi2 >> 8) >> 5) < Fb + kBlocks) {
                    const unsigned off2 = (unsigned)__shfl_sync(0xffffffffu, cur.x, c + 1);
                    const unsigned char* rec2 = stream + ((long long)off2 << 4);
                    {
                        const int pc = c + 1 + kEllAhead;
                        if (lane == (pc & 31)) prefetch_mine(pc < 32 ? cur : nxt);
                    }
                    switch ((max(nch, nch2) + 1) >> 1) {
                        case 1: ell_long2<MODE, 2>(rec, nch, rec2, nch2, lane, lw, s_log, lnl_local, re, rowid); break;
                        case 2: ell_long2<MODE, 4>(rec, nch, rec2, nch2, lane, lw, s_log, lnl_local, re, rowid); break;
                        case 3: ell_long2<MODE, 6>(rec, nch, rec2, nch2, lane, lw, s_log, lnl_local, re, rowid); break;
                        default: ell_long2<MODE, 8>(rec, nch, rec2, nch2, lane, lw, s_log, lnl_local, re, rowid); break;
                    }
                    __syncwarp();
                    ++c;
                    continue;
                }
            }
            switch ((nch + 1) >> 1) {
                case 1: ell_long<MODE, 2>(rec, nch, lane, lw, s_log, lnl_local, re, rowid); break;
                case 2: ell_long<MODE, 4>(rec, nch, lane, lw, s_log, lnl_local, re, rowid); break;
                case 3: ell_long<MODE, 6>(rec, nch, lane, lw, s_log, lnl_local, re, rowid); break;
                case 4: ell_long<MODE, 8>(rec, nch, lane, lw, s_log, lnl_local, re, rowid); break;
                default: ell_long_big<MODE>(rec, nch, lane, lw, s_log, lnl_local, re, rowid); break;
            }
            __syncwarp();          // the next read touches the same window words from other lanes
        }
        cur = nxt;
        nxt = after;
    }
    if (MODE != ELL_LNL) {
        __syncwarp();
        if (Fb >= 0) for (int blk = Fb; blk < Fb + kBlocks; ++blk) flush_block(blk);
    } else {
        lnl_local = group_sum<32>(lnl_local, 0xffffffffu);
        if (lane == 0) a.partials[blockIdx.x] = lnl_local;
    }
}

}  // namespace tsc
