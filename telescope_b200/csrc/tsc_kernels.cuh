// Device code of libtelescope_b200: Telescope's EM reassignment path on sm_100a.
//
// Kernel families over one HBM-resident shard (fp64 Q values, int32 locus indices, int64 row pointers per GPU):
//   * "rows" kernels (this file) -- one sub-warp of G lanes per read.  Simple; used for construction, per-entry outputs
//     (reassign data), the residual of the reassign sums, and as the in-library cross-check of the fast paths.
//   * "tiles" kernel (tsc_tiles.cuh) -- flat tiles of whole reads (<= 128 entries), one warp per tile, lane-consecutive
//     loads, one blocked segmented scan: posterior export (writes z), and the fused / log-likelihood passes over the reads
//     that are not in the clustered stream.
//   * "stream" kernels (tsc_ell.cuh) -- the per-iteration fused E+M step, the log-likelihood and the reassign sums over
//     the locus-clustered sliced-ELL copy of the ambiguous reads; (tsc_peer.cuh) the per-iteration tail with the exchange
//     between GPUs.
//
// Arithmetic follows the reference's operation order wherever it changes bits (compiled with -fmad=false):
//   n = Q * (pi*theta)  or  Q * pi          telescope/utils/model.py:718-720
//   z = n * recip0(sum_j n)                 telescope/utils/sparse_plus.py:16-22,52
//   c = (z * w) * Y                         telescope/utils/model.py:730-733
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace tsc {

struct EmState {          // lives in device memory, mirrored to pinned host memory for polling
    int iter;             // completed iterations (model.py:775 inum)
    int done;             // converged || reached_max -- kernels of later iterations exit immediately
    int converged;
    int pad;
    double lnl_prev;      // self.lnl inside the loop when use_likelihood (model.py:786-789)
    double lnl;
    double diff;
};

struct Consts {           // model.py:690-697,734,739
    double total_wt, ambig_wt, wmax, pi_prior_wt, theta_prior_wt, pi_denom, theta_denom;
};

struct Csr {              // one shard
    const long long* indptr;   // n_rows + 1
    const int* col;            // nnz (+ padding), internal (frequency-ordered) locus numbering
    const double* q;           // nnz (+ padding)
    long long n_rows;
};

__device__ __forceinline__ double recip0(double s) {
    // sparse_plus.py:16-22: 1/v with inf -> 0.  (1/denormal overflows to inf in numpy too, and is zeroed likewise.)
    double r = 1.0 / s;
    return isinf(r) ? 0.0 : r;
}

// log1p(x) for the log-likelihood (model.py:755-758).  The arguments are Q * pi * theta with Q = expm1(100 s / max):
// essentially always >= 2^53, where 1 + x == x in fp64 and log1p(x) == log(x).  For those a table-driven log is
// used: x = 2^e * m, m in [1,2); i = top 7 mantissa bits, c_i ~ 1/m (table), r = m*c_i - 1 with |r| <= 2^-8,
// log(x) = e*ln2 + (-log c_i) + (r - r^2/2 + ... - r^6/6)    [truncation < 2e-18, result >= 36]
// ~20 instructions instead of the ~100 of the library log1p.  Anything smaller (or not finite) takes log1p itself.
constexpr int kLogTab = 128;
struct LogTab { double c, l; };          // c_i, -log(c_i)

__global__ void k_log_table(LogTab* __restrict__ tab) {
    const int i = threadIdx.x;
    if (i >= kLogTab) return;
    const double c = 1.0 / (1.0 + (i + 0.5) / kLogTab);
    tab[i].c = c;
    tab[i].l = -log(c);
}

// branch-free core: log(x) for finite x >= 2^53 (any other bit pattern gives a harmless garbage value)
__device__ __forceinline__ double log_big_core(double x, const LogTab* __restrict__ s_tab) {
    const int hi = __double2hiint(x), lo = __double2loint(x);
    const int e = (hi >> 20) - 1023;
    const LogTab t = s_tab[(hi >> 13) & (kLogTab - 1)];
    const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
    const double r = fma(m, t.c, -1.0);
    double p = fma(r, -1.0 / 6.0, 1.0 / 5.0);
    p = fma(r, p, -1.0 / 4.0);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -1.0 / 2.0);
    p = fma(r * r, p, r);
    const double ed = (double)e;
    // ln2 split: the high part has 11 trailing zero bits, so e * hi is exact
    return fma(ed, 6.93147180369123816490e-01, t.l) + fma(ed, 1.90821492927058770002e-10, p);
}
__device__ __forceinline__ bool log_big_ok(double x) { return x >= 9007199254740992.0 && x <= 1.7e308; }   // 2^53 .. finite

__device__ __forceinline__ double log1p_big(double x, const LogTab* __restrict__ s_tab) {
    if (!log_big_ok(x)) return log1p(x);
    return log_big_core(x, s_tab);
}

template <int G>
struct GroupBits { static constexpr unsigned value = (1u << (G & 31)) - 1u; };
template <>
struct GroupBits<32> { static constexpr unsigned value = 0xffffffffu; };

template <int G>
__device__ __forceinline__ unsigned group_mask() {
    const unsigned lane = threadIdx.x & 31u;
    return GroupBits<G>::value << (lane & ~(unsigned)(G - 1));
}

template <int G>
__device__ __forceinline__ double group_sum(double v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, G);
    return v;
}
template <int G>
__device__ __forceinline__ double group_max(double v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(m, v, o, G));
    return v;
}
template <int G>
__device__ __forceinline__ int group_sum_int(int v, unsigned m) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(m, v, o, G);
    return v;
}

// deterministic block reduction (fixed tree), result valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* s_warp /* >= 32 doubles */) {
    v = group_sum<32>(v, 0xffffffffu);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s_warp[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? s_warp[threadIdx.x] : 0.0;
    if (w == 0) v = group_sum<32>(v, 0xffffffffu);
    return v;
}
__device__ __forceinline__ double block_max(double v, double* s_warp) {
    v = group_max<32>(v, 0xffffffffu);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) s_warp[w] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    v = (threadIdx.x < nw) ? s_warp[threadIdx.x] : 0.0;
    if (w == 0) v = group_max<32>(v, 0xffffffffu);
    return v;
}

// ---------------------------------------------------------------------------------------------------------------
// construction (model.py:635-700)
// ---------------------------------------------------------------------------------------------------------------

// Q = LUT[raw] (the LUT is numpy's expm1 evaluated on the host, so Q is bit-identical to model.py:653) and locus
// renumbering, one pass, 8 entries per thread.
__global__ void k_build_q(const uint16_t* __restrict__ raw, const int* __restrict__ col_in,
                          const double* __restrict__ lut, int lut_len, const int* __restrict__ perm,
                          double* __restrict__ q, int* __restrict__ col, long long nnz, int* __restrict__ bad) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; i < nnz; i += stride) {
        const int s = raw[i];
        if (s >= lut_len) { atomicOr(bad, 2); q[i] = 0.0; } else q[i] = __ldg(lut + s);
        col[i] = perm ? __ldg(perm + col_in[i]) : col_in[i];
    }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {   // splitmix64 finaliser
    x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL;
    x ^= x >> 27; x *= 0x94d049bb133111ebULL;
    x ^= x >> 31;
    return x;
}

// Per-locus entry count and an order-independent signature of the locus's column {(read, score)}: three 64-bit sums
// of independently mixed hashes, plus the number of positive scores.  Loci that agree in all five words (2 exact
// counts + 192 hash bits) are taken to hold identical columns; the reference gives such loci bit-identical pi/theta
// (its column sums run in read order), which reassign()'s exact tie test depends on.
__global__ void k_col_signature(const long long* __restrict__ indptr, long long n_rows, const int* __restrict__ col,
                                const uint16_t* __restrict__ raw, int n_cols, unsigned long long row_key0,
                                unsigned long long* __restrict__ sig /* 5*K: count, sum h1, sum h2, count of scores > 0, sum h3 */,
                                int* __restrict__ bad) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < n_rows; r += stride) {
        const unsigned long long rk = mix64(row_key0 + (unsigned long long)r);
        for (long long k = indptr[r]; k < indptr[r + 1]; ++k) {
            const int c = col[k];
            if (c < 0 || c >= n_cols) { atomicOr(bad, 1); continue; }
            const unsigned long long e = rk ^ ((unsigned long long)raw[k] * 0x9e3779b97f4a7c15ULL);
            atomicAdd(sig + c, 1ULL);
            atomicAdd(sig + n_cols + c, mix64(e));
            atomicAdd(sig + 2 * (size_t)n_cols + c, mix64(e ^ 0xd6e8feb86659fd93ULL));
            if (raw[k] != 0) atomicAdd(sig + 3 * (size_t)n_cols + c, 1ULL);
            atomicAdd(sig + 4 * (size_t)n_cols + c, mix64((e + 0x2545f4914f6cdd1dULL) * 0xff51afd7ed558ccdULL));
        }
    }
}

// w_i = max_j Q_ij (model.py:690); wy_i = w_i * Y_i with Y_i = [row has > 1 entries] (model.py:679);
// partial[0] += sum w, partial[1] += sum wy, partial[2] = max w; pisum0_j += Q_ij over unique reads (model.py:699).
__global__ void k_row_init(Csr a, int n_cols, double* __restrict__ wy, double* __restrict__ partial,
                           double* __restrict__ pisum0) {
    __shared__ double s_red[32];
    double sw = 0, swy = 0, mw = 0;
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (; r < a.n_rows; r += stride) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        double w = 0;                                   // sparse max: implicit zeros unless the row is full; Q >= 0
        for (long long k = s; k < e; ++k) w = fmax(w, a.q[k]);
        const bool amb = (e - s) > 1;
        wy[r] = amb ? w : 0.0;
        sw += w;
        if (amb) swy += w;
        mw = fmax(mw, w);
        if (!amb && e > s) {
            const int c = a.col[s];          // (may run before the loci have been validated: stay in bounds)
            if ((unsigned)c < (unsigned)n_cols) atomicAdd(pisum0 + c, a.q[s]);
        }
    }
    sw = block_sum(sw, s_red);
    if (threadIdx.x == 0) atomicAdd(partial + 0, sw);
    swy = block_sum(swy, s_red);
    if (threadIdx.x == 0) atomicAdd(partial + 1, swy);
    mw = block_max(mw, s_red);
    if (threadIdx.x == 0)   // non-negative doubles order like their bit patterns
        atomicMax(reinterpret_cast<unsigned long long*>(partial + 2), (unsigned long long)__double_as_longlong(mw));
}

__global__ void k_fill(double* __restrict__ p, long long n, double v) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

// tab[j] = a[j] * b[j]  (pi*theta, model.py:718)
__global__ void k_mul(const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}

// ---------------------------------------------------------------------------------------------------------------
// per-iteration K-length kernels
// ---------------------------------------------------------------------------------------------------------------

// thetasum_j = sum over the R accumulator replicas in fixed order; replicas are zeroed for the next iteration.
__global__ void k_reduce_replicas(double* __restrict__ acc, int K, int R, double* __restrict__ thetasum,
                                  const EmState* __restrict__ st) {
    if (st && st->done) return;
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    double s = 0;
    for (int r = 0; r < R; ++r) { s += acc[(size_t)r * K + j]; acc[(size_t)r * K + j] = 0.0; }
    thetasum[j] = s;
}

struct UpdateArgs {
    const double* thetasum;   // global (all-reduced) per-locus sums of z*w*Y
    const double* pisum0;
    const Consts* c;
    double *pi, *theta, *pt;                  // current parameters; overwritten with the new estimates
    double *pi_prev, *theta_prev, *pt_prev;   // parameters the last E-step used (the stored z of model.py:795)
    double *pi_init, *theta_init;
    EmState* st;
    double* diffs;
    int K, max_iter, use_lnl;
    double eps;
    const int* rep;           // rep[j] = first locus whose column is identical to j's (nullptr: all distinct)
};

// mstep tail (model.py:734-742), diff_est (model.py:781) and the loop control of model.py:788-796.  One block:
// K <= a few 10^5 and the reduction order must not depend on the launch shape.
__global__ void __launch_bounds__(1024) k_update(UpdateArgs a) {
    __shared__ double s_red[32];
    EmState* st = a.st;
    if (st->done) return;
    const int iter = st->iter;
    const Consts c = *a.c;
    double local = 0;
    for (int j = threadIdx.x; j < a.K; j += blockDim.x) {
        const double ts = a.thetasum[a.rep ? a.rep[j] : j];    // identical loci share one sum -> exact ties survive
        const double th = (ts + c.theta_prior_wt) / c.theta_denom;
        const double pisum = a.pisum0[j] + ts;
        const double pn = (pisum + c.pi_prior_wt) / c.pi_denom;
        const double po = a.pi[j];
        local += fabs(pn - po);
        a.pi_prev[j] = po; a.theta_prev[j] = a.theta[j]; a.pt_prev[j] = a.pt[j];
        a.pi[j] = pn; a.theta[j] = th; a.pt[j] = pn * th;
        if (iter == 0) { a.pi_init[j] = pn; a.theta_init[j] = th; }
    }
    const double diff = block_sum(local, s_red);
    if (threadIdx.x == 0) {
        a.diffs[iter] = diff;
        st->diff = diff;
        st->iter = iter + 1;
        if (!a.use_lnl) {
            const int conv = diff < a.eps;
            st->converged = conv;
            st->done = conv || (iter + 1 >= a.max_iter);
        }
    }
}

// use_likelihood loop control (model.py:785-789): lnl is the all-reduced log-likelihood of this iteration
__global__ void k_lnl_control(EmState* st, const double* lnl, double* lnls, double eps, int max_iter) {
    if (st->done) return;
    const double v = *lnl;
    lnls[st->iter - 1] = v;
    const int conv = fabs(v - st->lnl_prev) < eps;
    st->lnl_prev = v;
    st->lnl = v;
    st->converged = conv;
    st->done = conv || (st->iter >= max_iter);
}

// ---------------------------------------------------------------------------------------------------------------
// "rows" kernels: G lanes per read
// ---------------------------------------------------------------------------------------------------------------

// One read: lanes stride over its entries, n = Q * tab[col]; returns the group-wide sum.  The first chunk's values
// stay in registers (n0, c0, k0) so that reads with <= G entries are touched once.
template <int G>
__device__ __forceinline__ double row_numerators(const Csr& a, long long s, long long e, const double* __restrict__ tab,
                                                 unsigned m, int lane, double& n0, int& c0) {
    double sum = 0;
    n0 = 0; c0 = 0;
    long long k = s + lane;
    if (k < e) { c0 = a.col[k]; n0 = a.q[k] * __ldg(tab + c0); sum = n0; k += G; }
    for (; k < e; k += G) sum += a.q[k] * __ldg(tab + a.col[k]);
    return group_sum<G>(sum, m);
}

// Fused E-step + M-step accumulation for one EM iteration (model.py:718-722 + 730-733): nothing but the K-length
// accumulator is written.  pt = pi*theta.  Unique reads (Y=0) contribute nothing here (they are in pisum0).
template <int G>
__global__ void __launch_bounds__(512) k_fused_rows(Csr a, const double* __restrict__ wy, const double* __restrict__ pt,
                                                    double* __restrict__ acc, int K, int R, const EmState* __restrict__ st) {
    if (st && st->done) return;
    double* my = acc + (size_t)(blockIdx.x % R) * K;
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        if (e - s < 2) continue;
        double n0; int c0;
        const double sum = row_numerators<G>(a, s, e, pt, m, lane, n0, c0);
        const double rr = recip0(sum);
        const double w = wy[r];
        long long k = s + lane;
        if (k < e) { const double c = (n0 * rr) * w; if (c != 0.0) atomicAdd(my + c0, c); k += G; }
        for (; k < e; k += G) {
            const int cc = a.col[k];
            const double c = ((a.q[k] * __ldg(pt + cc)) * rr) * w;
            if (c != 0.0) atomicAdd(my + cc, c);
        }
    }
}

// E-step alone (model.py:702-722): z for every stored entry.  tab_amb = pi*theta, tab_uni = pi.
template <int G>
__global__ void __launch_bounds__(512) k_estep_rows(Csr a, const double* __restrict__ tab_amb, const double* __restrict__ tab_uni,
                                                    double* __restrict__ z) {
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const double* tab = (e - s > 1) ? tab_amb : tab_uni;
        double n0; int c0;
        const double sum = row_numerators<G>(a, s, e, tab, m, lane, n0, c0);
        const double rr = recip0(sum);
        long long k = s + lane;
        if (k < e) { z[k] = n0 * rr; k += G; }
        for (; k < e; k += G) z[k] = (a.q[k] * __ldg(tab + a.col[k])) * rr;
    }
}

// M-step accumulation from a caller-supplied z (model.py:730-733), for the tsc_mstep entry point
template <int G>
__global__ void __launch_bounds__(512) k_mstep_rows(Csr a, const double* __restrict__ wy, const double* __restrict__ z,
                                                    double* __restrict__ acc, int K, int R) {
    double* my = acc + (size_t)(blockIdx.x % R) * K;
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        if (e - s < 2) continue;
        const double w = wy[r];
        for (long long k = s + lane; k < e; k += G) {
            const double c = z[k] * w;
            if (c != 0.0) atomicAdd(my + a.col[k], c);
        }
    }
}

// log-likelihood (model.py:744-760): sum_ij z_ij * log1p(Q_ij * pi_j * theta_j^Y_i).
// z is either given (zin) or regenerated from the E-step tables (zamb/zuni).  Per-block partials, reduced in
// fixed order by k_sum_partials.
template <int G>
__global__ void __launch_bounds__(512) k_lnl_rows(Csr a, const double* __restrict__ zin,
                                                  const double* __restrict__ zamb, const double* __restrict__ zuni,
                                                  const double* __restrict__ inner_amb, const double* __restrict__ inner_uni,
                                                  double* __restrict__ partials, const EmState* __restrict__ st) {
    __shared__ double s_red[32];
    if (st && st->done) return;
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    double local = 0;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const bool amb = (e - s) > 1;
        const double* in = amb ? inner_amb : inner_uni;
        if (zin) {
            for (long long k = s + lane; k < e; k += G) {
                const double zz = zin[k];
                if (zz != 0.0) local += zz * log1p(a.q[k] * __ldg(in + a.col[k]));
            }
        } else {
            const double* tab = amb ? zamb : zuni;
            double n0; int c0;
            const double sum = row_numerators<G>(a, s, e, tab, m, lane, n0, c0);
            const double rr = recip0(sum);
            for (long long k = s + lane; k < e; k += G) {
                const int cc = a.col[k];
                const double q = a.q[k];
                const double zz = (q * __ldg(tab + cc)) * rr;
                if (zz != 0.0) local += zz * log1p(q * __ldg(in + cc));
            }
        }
    }
    local = block_sum(local, s_red);
    if (threadIdx.x == 0) partials[blockIdx.x] = local;
}

__global__ void __launch_bounds__(1024) k_sum_partials(const double* __restrict__ partials, int n, double* __restrict__ out) {
    __shared__ double s_red[32];
    double v = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) v += partials[i];
    v = block_sum(v, s_red);
    if (threadIdx.x == 0) *out = v;
}

// reassign (model.py:808-865, sparse_plus.py:99-165).  For each read: z from the E-step tables, best hits are the
// stored (non-zero) entries equal to the row maximum (exact comparison, as sparse_plus.py:125).
// Outputs, all optional: nbest per read, per-locus column sums (atomics), per-entry assignment values.
struct ReassignArgs {
    int method;
    double thresh;
    const int* picks;     // per read, TSC_CHOOSE only
    int* nbest;           // per read
    double* colsum;       // K
    double* data;         // nnz
    const int* rowid;     // optional: nbest is indexed by rowid[read] (residual CSR -> the shard's reads)
};

template <int G>
__global__ void __launch_bounds__(512) k_reassign_rows(Csr a, const double* __restrict__ tab_amb, const double* __restrict__ tab_uni,
                                                       ReassignArgs g) {
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const bool amb = (e - s) > 1;
        const double* tab = amb ? tab_amb : tab_uni;
        double n0; int c0;
        const double sum = row_numerators<G>(a, s, e, tab, m, lane, n0, c0);
        const double rr = recip0(sum);
        // pass 2: row maximum of z and, for conf, the sum of the surviving entries
        double zmax = 0, kept = 0;
        for (long long k = s + lane; k < e; k += G) {
            const double zz = (a.q[k] * __ldg(tab + a.col[k])) * rr;
            zmax = fmax(zmax, zz);
            if (zz >= g.thresh) kept += zz;
        }
        zmax = group_max<G>(zmax, m);
        kept = group_sum<G>(kept, m);
        // pass 3: number of best hits
        int nb = 0;
        for (long long k = s + lane; k < e; k += G) {
            const double zz = (a.q[k] * __ldg(tab + a.col[k])) * rr;
            nb += (zz == zmax && zz != 0.0);
        }
        nb = group_sum_int<G>(nb, m);
        if (g.nbest && lane == 0) g.nbest[g.rowid ? g.rowid[r] : r] = nb;
        if (!g.colsum && !g.data) continue;
        const int pick = (g.method == 1 && g.picks && nb > 1) ? g.picks[r] : 0;
        const double rkept = recip0(kept);
        const double ravg = (nb > 0) ? 1.0 / (double)nb : 0.0;
        int seen = 0;   // best hits before this chunk (row order)
        const long long e_round = s + ((e - s + G - 1) / G) * G;
        for (long long k = s + lane; k < e_round; k += G) {
            const bool act = k < e;
            int cc = 0; double zz = 0;
            if (act) { cc = a.col[k]; zz = (a.q[k] * __ldg(tab + cc)) * rr; }
            const bool best = act && zz == zmax && zz != 0.0;
            double val = 0;
            switch (g.method) {
                case 0: val = (best && nb == 1) ? 1.0 : 0.0; break;
                case 1: {
                    const unsigned bal = (__ballot_sync(m, best) >> ((threadIdx.x & 31u) & ~(unsigned)(G - 1))) & GroupBits<G>::value;
                    const int rank = seen + __popc(bal & ((1u << lane) - 1u));
                    val = (best && (nb == 1 || rank == pick)) ? 1.0 : 0.0;
                    seen += __popc(bal);
                } break;
                case 2: val = best ? ravg : 0.0; break;
                case 3: val = (act && zz >= g.thresh) ? zz * rkept : 0.0; break;
                case 4: val = (act && !amb) ? ceil(zz) : 0.0; break;
                default: val = (act && zz > 0.0) ? 1.0 : 0.0; break;
            }
            if (act) {
                if (g.data) g.data[k] = val;
                if (g.colsum && val != 0.0) atomicAdd(g.colsum + cc, val);
            }
        }
    }
}

// Everything Telescope.output_report needs (model.py:432-458) in ONE pass over the shard: for each read the final
// posterior z_f (E-step tables) and the initial one z_i = Q.norm(1) are formed side by side, and six per-locus sums
// are accumulated:
//   out[0] final_conf     reassign('conf', thresh)            out[3] init_best      reassign('exclude', initial)
//   out[1] (not here: init_aligned = stored entries with a positive score per locus, known since construction)
//   out[2] unique_count   reassign('unique')                   out[4] init_best_avg  reassign('average', initial)
//   out[5] the counts file's reassign(final_method, thresh); for 'choose' only reads with a single best hit are
//          added here -- the tie reads need the host's RNG draws (nbest_final) and a second, cheap pass.
// nbest_init feeds the draws of init_best_random = reassign('choose', initial).
struct ReportArgs {
    double thresh;
    int final_method;
    int* nbest_init;
    int* nbest_final;
    double* out;          // 6 * K
    int K;
};

template <int G>
__global__ void __launch_bounds__(512) k_report_rows(Csr a, const double* __restrict__ tab_amb, const double* __restrict__ tab_uni,
                                                     ReportArgs g) {
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    const int K = g.K;
    for (; r < a.n_rows; r += ngrp) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const bool amb = (e - s) > 1;
        const double* tab = amb ? tab_amb : tab_uni;
        // pass 1: both row sums
        double sum_f = 0, sum_i = 0;
        for (long long k = s + lane; k < e; k += G) { const double q = a.q[k]; sum_f += q * __ldg(tab + a.col[k]); sum_i += q; }
        sum_f = group_sum<G>(sum_f, m);
        sum_i = group_sum<G>(sum_i, m);
        const double rr_f = recip0(sum_f), rr_i = recip0(sum_i);
        // pass 2: row maxima and the conf survivors
        double max_f = 0, max_i = 0, kept = 0;
        for (long long k = s + lane; k < e; k += G) {
            const double q = a.q[k];
            const double zf = (q * __ldg(tab + a.col[k])) * rr_f, zi = q * rr_i;
            max_f = fmax(max_f, zf);
            max_i = fmax(max_i, zi);
            if (zf >= g.thresh) kept += zf;
        }
        max_f = group_max<G>(max_f, m);
        max_i = group_max<G>(max_i, m);
        kept = group_sum<G>(kept, m);
        // pass 3: best-hit counts
        int nb_f = 0, nb_i = 0;
        for (long long k = s + lane; k < e; k += G) {
            const double q = a.q[k];
            const double zf = (q * __ldg(tab + a.col[k])) * rr_f, zi = q * rr_i;
            nb_f += (zf == max_f && zf != 0.0);
            nb_i += (zi == max_i && zi != 0.0);
        }
        nb_f = group_sum_int<G>(nb_f, m);
        nb_i = group_sum_int<G>(nb_i, m);
        if (lane == 0) {
            if (g.nbest_init) g.nbest_init[r] = nb_i;
            if (g.nbest_final) g.nbest_final[r] = nb_f;
        }
        const double rkept = recip0(kept);
        const double avg_f = nb_f > 0 ? 1.0 / (double)nb_f : 0.0, avg_i = nb_i > 0 ? 1.0 / (double)nb_i : 0.0;
        // pass 4: contributions
        for (long long k = s + lane; k < e; k += G) {
            const int c = a.col[k];
            const double q = a.q[k];
            const double zf = (q * __ldg(tab + c)) * rr_f, zi = q * rr_i;
            const bool best_f = (zf == max_f && zf != 0.0), best_i = (zi == max_i && zi != 0.0);
            const double conf = (zf >= g.thresh) ? zf * rkept : 0.0;
            const double uniq = amb ? 0.0 : ceil(zf);
            if (conf != 0.0) atomicAdd(g.out + c, conf);
            if (uniq != 0.0) atomicAdd(g.out + 2 * (size_t)K + c, uniq);
            if (best_i && nb_i == 1) atomicAdd(g.out + 3 * (size_t)K + c, 1.0);
            if (best_i) atomicAdd(g.out + 4 * (size_t)K + c, avg_i);
            double fin = 0;
            switch (g.final_method) {
                case 0: case 1: fin = (best_f && nb_f == 1) ? 1.0 : 0.0; break;
                case 2: fin = best_f ? avg_f : 0.0; break;
                case 3: fin = conf; break;
                case 4: fin = uniq; break;
                default: fin = (zf > 0.0) ? 1.0 : 0.0; break;
            }
            if (fin != 0.0) atomicAdd(g.out + 5 * (size_t)K + c, fin);
        }
    }
}

// reassign('choose') for the reads with several best hits only: adds 1 to the locus of the picks[r]-th best hit
template <int G>
__global__ void __launch_bounds__(512) k_choose_ties_rows(Csr a, const double* __restrict__ tab_amb, const double* __restrict__ tab_uni,
                                                          const int* __restrict__ nbest, const int* __restrict__ picks,
                                                          double* __restrict__ colsum) {
    const unsigned m = group_mask<G>();
    const int lane = threadIdx.x & (G - 1);
    long long r = (blockIdx.x * (long long)blockDim.x + threadIdx.x) / G;
    const long long ngrp = (long long)gridDim.x * blockDim.x / G;
    for (; r < a.n_rows; r += ngrp) {
        if (nbest[r] < 2) continue;                    // group-uniform
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const double* tab = (e - s > 1) ? tab_amb : tab_uni;
        double n0; int c0;
        const double sum = row_numerators<G>(a, s, e, tab, m, lane, n0, c0);
        const double rr = recip0(sum);
        double zmax = 0;
        for (long long k = s + lane; k < e; k += G) zmax = fmax(zmax, (a.q[k] * __ldg(tab + a.col[k])) * rr);
        zmax = group_max<G>(zmax, m);
        const int pick = picks[r];
        int seen = 0;
        const long long e_round = s + ((e - s + G - 1) / G) * G;
        for (long long k = s + lane; k < e_round; k += G) {
            const bool act = k < e;
            int cc = 0; double zz = 0;
            if (act) { cc = a.col[k]; zz = (a.q[k] * __ldg(tab + cc)) * rr; }
            const bool best = act && zz == zmax && zz != 0.0;
            const unsigned bal = (__ballot_sync(m, best) >> ((threadIdx.x & 31u) & ~(unsigned)(G - 1))) & GroupBits<G>::value;
            if (best && seen + __popc(bal & ((1u << lane) - 1u)) == pick) atomicAdd(colsum + cc, 1.0);
            seen += __popc(bal);
        }
    }
}

// w per read for the API (model.py:690): wy for ambiguous reads, the single Q for unique ones; Y likewise
__global__ void k_row_info(Csr a, const double* __restrict__ wy, uint8_t* __restrict__ y, double* __restrict__ w) {
    long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    for (; r < a.n_rows; r += (long long)gridDim.x * blockDim.x) {
        const long long s = a.indptr[r], e = a.indptr[r + 1];
        const bool amb = (e - s) > 1;
        if (y) y[r] = amb ? 1 : 0;
        if (w) w[r] = amb ? wy[r] : (e > s ? fmax(a.q[s], 0.0) : 0.0);
    }
}

}  // namespace tsc
