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
// The slice records form one byte stream in HBM; a warp takes segments of 32 consecutive records and pulls every record
// with ONE 1-D TMA bulk copy (cp.async.bulk + mbarrier complete_tx) into a 12 KB shared-memory ring, several records
// ahead of the arithmetic, so 7 one-warp CTAs per SM keep enough bytes in flight for the HBM roofline with no load
// instruction on the global path.
//
// Record (16-byte aligned, 144 + 288*T bytes):
//   int32 lo, T, hi, n_reads | double wy[16] | uint8 wcol[T][32] | double q[T][32]
// wcol = locus & 127, the entry's row in the window (blocks of 32 loci live in slot (locus >> 5) & 3), or 128 for an
// empty slot: the dummy row, with q = 0 and pi*theta = 0, so empty slots add an exact +0.0 to a word nobody reads.
// With one byte per locus the stream moves ~9.4 B per entry where canonical CSR moves 12.
//
// Reads that do not fit a slice (fewer than 2 or more than 2*kEllTMax entries, a locus span above kEllSpan, or
// non-increasing loci) are not in the stream: unique reads add nothing to the M-step sums (model.py:730-733; they
// enter pi through pisum0, model.py:699,738) and the rest go through the flat-tile kernel on a compact residual CSR.
#pragma once
#include "tsc_kernels.cuh"

namespace tsc {

constexpr int kEllReads = 16;                 // reads per slice
constexpr int kEllTMax = 24;                  // steps per slice; a read holds at most 2*kEllTMax entries
constexpr int kEllWin = 128;                  // loci in a warp's window (4 blocks of 32)
constexpr int kEllSpan = 96;                  // max (locus - slice lo) inside a slice; lo % 32 + span < kEllWin
constexpr int kEllHdr = 16 + kEllReads * 8;   // header + wy
constexpr int kEllRing = 12288;               // bytes of record ring per warp (a record is 144 + 288*T <= 7056 bytes)
constexpr int kEllQueue = 8;                  // records in flight per warp (mbarrier slots)
constexpr int kEllLenBits = 6;                // sort key = first locus << 6 | snake(length)
constexpr size_t kEllSmem = kEllRing + sizeof(double) * ((kEllWin + 1) * kEllReads + kEllWin + 8) + 8 * kEllQueue + 4 * kEllQueue;
static_assert(2 * kEllTMax < (1 << kEllLenBits), "length must fit the key");
static_assert(kEllTMax == 24, "k_ell_fused dispatches bodies of 4..24 steps");
static_assert(kEllHdr + 288 * kEllTMax <= kEllRing, "the largest record must fit the ring");

__host__ __device__ inline int ell_record_bytes(int T) { return kEllHdr + 288 * T; }

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

// Per read: the sort key of a slice candidate (first locus, then length in snake order so that neighbouring keys
// hold reads of similar length), or -1; candidates are counted per key.
__global__ void k_ell_classify(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ col, int n_keys,
                               int* __restrict__ key, unsigned* __restrict__ hist) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const long long s = ip[r], e = ip[r + 1];
        const long long len = e - s;
        int k = -1;
        if (len >= 2 && len <= 2 * kEllTMax) {
            const int first = col[s];
            bool ok = true;
            int prev = first;
            for (long long p = s + 1; p < e; ++p) { const int c = col[p]; ok = ok && (c > prev); prev = c; }
            if (ok && prev - first <= kEllSpan) {
                const int lk = (first & 1) ? ((1 << kEllLenBits) - 1 - (int)len) : (int)len;
                k = (first << kEllLenBits) | lk;
                if (k >= n_keys) k = -1;      // cannot happen for first < n_cols; keeps the histogram in bounds
            }
        }
        key[r] = k;
        if (k >= 0) atomicAdd(hist + k, 1u);
    }
}

// Counting-sort scatter: candidates land in key order (order inside a key is whatever the atomics give; all reads
// of a key have the same first locus and length, so only the summation order inside a slice depends on it).
__global__ void k_ell_scatter(const int* __restrict__ key, long long n_rows, const long long* __restrict__ bin_start,
                              unsigned* __restrict__ cursor, int* __restrict__ sorted) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const int k = key[r];
        if (k < 0) continue;
        const long long pos = bin_start[k] + atomicAdd(cursor + k, 1u);
        sorted[pos] = (int)r;
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
        hdr[s] = make_int4(lo, T, hi, kept);
        rec_bytes[s] = ell_record_bytes(T);
    }
}

// One warp per slice writes its record.
__global__ void __launch_bounds__(256) k_ell_fill(const long long* __restrict__ ip, const int* __restrict__ col,
                                                  const double* __restrict__ q, const double* __restrict__ wy,
                                                  const int* __restrict__ sorted, long long n_cand, long long n_slices,
                                                  const int4* __restrict__ hdr, const long long* __restrict__ rec_off,
                                                  unsigned char* __restrict__ stream) {
    const int lane = threadIdx.x & 31, r = lane & 15, h = lane >> 4;
    long long s = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 5;
    for (; s < n_slices; s += stride) {
        const int4 hd = hdr[s];
        unsigned char* rec = stream + rec_off[s];
        if (lane == 0) *reinterpret_cast<int4*>(rec) = hd;
        const long long p = s * kEllReads + r;
        const int read = (p < n_cand) ? sorted[p] : -1;
        long long b = 0;
        int len = 0;
        if (read >= 0) { b = ip[read]; len = (int)(ip[read + 1] - b); }
        if (h == 0) reinterpret_cast<double*>(rec + 16)[r] = (read >= 0) ? wy[read] : 0.0;
        const int T = hd.y;
        unsigned char* dc = rec + kEllHdr;
        double* qq = reinterpret_cast<double*>(rec + kEllHdr + 32 * T);
        for (int t = 0; t < T; ++t) {
            const int k = 2 * t + h;
            double qv = 0.0;
            int d = kEllWin;                                          // empty slot: the dummy row
            if (k < len) { qv = q[b + k]; d = col[b + k] & (kEllWin - 1); }
            dc[t * 32 + lane] = (unsigned char)d;
            qq[t * 32 + lane] = qv;
        }
    }
}

// Residual CSR = ambiguous reads that are not in the stream.  flag/len per read, then (after the scans) the copy.
__global__ void k_res_flags(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ key,
                            int* __restrict__ flag, int* __restrict__ rlen) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const long long len = ip[r + 1] - ip[r];
        const bool res = (len >= 2) && (key == nullptr || key[r] < 0);
        flag[r] = res ? 1 : 0;
        rlen[r] = res ? (int)min(len, (long long)0x7fffffff) : 0;
    }
}

__global__ void __launch_bounds__(256) k_res_copy(const long long* __restrict__ ip, long long n_rows, const int* __restrict__ col,
                                                  const double* __restrict__ q, const double* __restrict__ wy,
                                                  const int* __restrict__ flag, const long long* __restrict__ row_pos,
                                                  const long long* __restrict__ ent_pos, long long* __restrict__ ip_out,
                                                  int* __restrict__ col_out, double* __restrict__ q_out,
                                                  double* __restrict__ wy_out) {
    const int lane = threadIdx.x & 7;
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 3;
    const long long stride = ((long long)gridDim.x * blockDim.x) >> 3;
    for (; r < n_rows; r += stride) {
        if (!flag[r]) continue;
        const long long b = ip[r], e = ip[r + 1], o = ent_pos[r], rp = row_pos[r];
        if (lane == 0) { ip_out[rp] = o; wy_out[rp] = wy[r]; }
        for (long long k = b + lane; k < e; k += 8) { col_out[o + (k - b)] = col[k]; q_out[o + (k - b)] = q[k]; }
    }
}

// ---------------------------------------------------------------------------------------------------------------
// the per-iteration kernel
// ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned ell_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ell_mbar_init(unsigned bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void ell_mbar_expect_tx(unsigned bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ell_bulk_load(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void ell_mbar_wait(unsigned bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "ELL_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra ELL_DONE;\n\t"
        "bra ELL_WAIT;\n\t"
        "ELL_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}

struct EllArgs {
    const unsigned char* stream;
    const long long* rec_off;     // n_slices + 1 byte offsets of the slice records
    long long n_slices;
    const double* pt;             // pi*theta
    double* acc;                  // R replicas of K doubles
    int K, R;
    const EmState* st;            // nullptr = always run
};

// One slice record, TM = T rounded up to a multiple of 4: straight-line code, every load of the slice can be in flight
// at once; only the last three steps are conditional.
template <int TM>
__device__ __forceinline__ void ell_body(const unsigned char* rec, int T, int lane, const double* s_pt,
                                         unsigned char* accb /* s_acc + 8 * (lane & 15) */) {
    const double w_mine = reinterpret_cast<const double*>(rec + 16)[lane & 15];
    const unsigned char* cp = rec + kEllHdr + lane;
    const double* qp = reinterpret_cast<const double*>(rec + kEllHdr + 32 * T) + lane;
    double n[TM];
    unsigned ao[TM];              // byte offset of the entry's accumulator row
    double s0 = 0.0, s1 = 0.0;
    // ---- pass 1: numerators n = Q * (pi*theta)[locus], private row sum
#pragma unroll
    for (int t = 0; t < TM; ++t) {
        n[t] = 0.0;
        ao[t] = kEllWin * kEllReads * 8;            // dummy row
        if (t < TM - 3 || t < T) {
            const unsigned jw = cp[32 * t];
            ao[t] = jw * (kEllReads * 8);
            n[t] = qp[32 * t] * s_pt[jw];
            if (t & 1) s1 += n[t]; else s0 += n[t];
        }
    }
    double sum = s0 + s1;
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    // (w*Y) * recip0(total), the same expression as the tile kernel; empty read slots have w = 0
    const double g = (w_mine != 0.0) ? w_mine * recip0(sum) : 0.0;
    // ---- pass 2: c = n * g into this read slot's private accumulator column.  Loci are unique within a read, the two
    // lanes of a read hold different entries, and empty slots point at the dummy row, so no two real updates of a
    // slice share an address: all loads may precede all stores.
#pragma unroll
    for (int t = 0; t < TM; ++t) n[t] = *reinterpret_cast<const double*>(accb + ao[t]) + n[t] * g;
#pragma unroll
    for (int t = 0; t < TM; ++t) *reinterpret_cast<double*>(accb + ao[t]) = n[t];
}

__global__ void __launch_bounds__(32) k_ell_fused(const EllArgs a) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    if (a.st && a.st->done) return;
    unsigned char* ring = s_raw;
    double* s_acc = reinterpret_cast<double*>(s_raw + kEllRing);          // [kEllWin + 1][kEllReads]; last row = dummy
    double* s_pt = s_acc + (kEllWin + 1) * kEllReads;                    // [kEllWin + 8]; [kEllWin] = 0 for empty slots
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_pt + kEllWin + 8);
    int* s_rpos = reinterpret_cast<int*>(s_bar + kEllQueue);
    const int lane = threadIdx.x;
    const int K = a.K;
    const double* __restrict__ pt = a.pt;
    double* my = a.acc + (size_t)(blockIdx.x % a.R) * K;
    const unsigned ring_u32 = ell_smem_u32(ring), bar_u32 = ell_smem_u32(s_bar);
    unsigned char* accb = reinterpret_cast<unsigned char*>(s_acc) + 8 * (lane & 15);

    for (int i = lane; i < (kEllWin + 1) * kEllReads + kEllWin + 8; i += 32) s_acc[i] = 0.0;     // accumulators and s_pt
    if (lane == 0) {
        for (int i = 0; i < kEllQueue; ++i) ell_mbar_init(bar_u32 + 8 * i, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncwarp();

    const int nw = gridDim.x;
    const long long n_slices = a.n_slices;
    const int n_seg = (int)((n_slices + 31) >> 5);            // a segment = 32 consecutive records
    // lane l of a segment's register set holds record 32*seg + l: byte offset and size
    auto load_seg = [&](int seg, long long& off, int& sz, int& count) {
        off = 0; sz = 0; count = 0;
        if (seg < n_seg) {
            const long long idx = ((long long)seg << 5) + lane;
            count = (int)min(32LL, n_slices - ((long long)seg << 5));
            if (idx < n_slices) { off = a.rec_off[idx]; sz = (int)(a.rec_off[idx + 1] - off); }
        }
    };
    // ---- producer state (tracked by every lane, issued by lane 0): the next record to request
    int prec = 0, pcount, ncount;
    long long poff, noff;
    int psz, nsz;
    load_seg(blockIdx.x, poff, psz, pcount);
    load_seg(blockIdx.x + nw, noff, nsz, ncount);
    int pnext_seg = blockIdx.x + 2 * nw;
    unsigned issued = 0, consumed = 0;                // records, per warp
    int head = 0, tail = 0;                           // ring allocator: in-flight records occupy [tail .. head) in FIFO order
    int Fb = -1;                                      // first block (32 loci) of the window; -1 = empty

    auto flush_block = [&](int b) {
        const int row = ((b & 3) * 32 + lane) * kEllReads;
        double s = 0.0;
#pragma unroll
        for (int c = 0; c < kEllReads; ++c) s += s_acc[row + ((c + lane) & 15)];
#pragma unroll
        for (int c = 0; c < kEllReads; ++c) s_acc[row + ((c + lane) & 15)] = 0.0;
        const int j = b * 32 + lane;
        if (j < K && s != 0.0) atomicAdd(my + j, s);
    };

    for (int seg = blockIdx.x; seg < n_seg; seg += nw) {
        const int nrec = (int)min(32LL, n_slices - ((long long)seg << 5));
        for (int c = 0; c < nrec; ++c) {
            // ---- everything before the current record is free; keep the ring full
            __syncwarp();
            if (issued == consumed) { head = 0; tail = 0; }
            else tail = s_rpos[consumed % kEllQueue];
            while (issued - consumed < kEllQueue - 1 && pcount > 0) {
                const int S = __shfl_sync(0xffffffffu, psz, prec);
                int place;
                if (issued == consumed) place = 0;
                else if (head > tail) { place = (kEllRing - head >= S) ? head : ((S <= tail) ? 0 : -1); }
                else place = (tail - head >= S) ? head : -1;
                if (place < 0) break;
                const long long O = __shfl_sync(0xffffffffu, poff, prec);
                if (lane == 0) {
                    const unsigned slot = issued % kEllQueue;
                    s_rpos[slot] = place;
                    ell_mbar_expect_tx(bar_u32 + 8 * slot, (unsigned)S);
                    ell_bulk_load(ring_u32 + place, a.stream + O, (unsigned)S, bar_u32 + 8 * slot);
                }
                head = place + S;
                ++issued;
                if (++prec == pcount) {
                    poff = noff; psz = nsz; pcount = ncount; prec = 0;
                    load_seg(pnext_seg, noff, nsz, ncount);
                    pnext_seg += nw;
                }
            }
            // ---- wait for the current record (an empty ring put it at 0)
            ell_mbar_wait(bar_u32 + 8 * (consumed % kEllQueue), (consumed / kEllQueue) & 1u);
            const unsigned char* rec = ring + tail;
            const int4 hd = *reinterpret_cast<const int4*>(rec);
            const int lo = hd.x, T = hd.y, hi = hd.z;

            // ---- window: blocks [Fb, Fb+4) of 32 loci; the stream is sorted by lo, so it only moves forward
            const int lb = lo >> 5;
            if (Fb < 0 || (hi >> 5) >= Fb + 4) {
                __syncwarp();
                int first_new = lb;
                if (Fb >= 0) {
                    const int e = min(lb, Fb + 4);
                    for (int b = Fb; b < e; ++b) flush_block(b);
                    first_new = max(lb, Fb + 4);
                }
                for (int b = first_new; b < lb + 4; ++b) {
                    const int j = b * 32 + lane;
                    s_pt[(b & 3) * 32 + lane] = (j < K) ? __ldg(pt + j) : 0.0;
                }
                Fb = lb;
                __syncwarp();
            }
            switch ((T + 3) >> 2) {
                case 1: ell_body<4>(rec, T, lane, s_pt, accb); break;
                case 2: ell_body<8>(rec, T, lane, s_pt, accb); break;
                case 3: ell_body<12>(rec, T, lane, s_pt, accb); break;
                case 4: ell_body<16>(rec, T, lane, s_pt, accb); break;
                case 5: ell_body<20>(rec, T, lane, s_pt, accb); break;
                default: ell_body<24>(rec, T, lane, s_pt, accb); break;
            }
            ++consumed;
        }
        // ---- segment done: hand the window to the global accumulator
        __syncwarp();
        if (Fb >= 0) for (int b = Fb; b < Fb + 4; ++b) flush_block(b);
        Fb = -1;
    }
}

}  // namespace tsc
