// Peer-memory transport of one node and the fused per-iteration tail.
//
// The only exchange of the EM path is the K-length vector of per-locus M-step sums, once per iteration
// (model.py:730-733 computes it over all reads; here every GPU holds a block of reads).  K doubles are 240 KB at
// K = 30 000: pure latency.  Instead of NCCL (a collective launch per iteration and a multi-second
// ncclCommInitRank per model) every GPU owns an EXCHANGE BUFFER that all ranks of the node map -- directly when one
// process drives several GPUs, through CUDA IPC when there is one process per GPU -- and the exchange happens inside
// the kernel that needs the result:
//
//   k_tail, one launch per iteration and GPU, ~K/1024 blocks:
//     1. sum this GPU's accumulator replicas for the block's loci (and zero them for the next iteration)
//     2. PUSH the partial sums into every rank's inbox[parity][my rank] (remote stores over NVLink), fence, then
//        raise flag[my rank][block] = epoch on every rank (release, system scope)
//     3. wait until every (rank, block) flag of the own buffer has reached this epoch (acquire, system scope)
//     4. add the ranks' partial sums in FIXED RANK ORDER -> the global sum, bit-identical on every rank
//     5. MAP update of pi/theta (model.py:734-742), |pi' - pi| per block; the last block to finish adds the block
//        partials in block order (model.py:781) and takes the loop decision (model.py:788-796)
//   It replaces k_reduce_replicas + ncclAllReduce + k_update (3 launches and ~135 us per iteration in round 1).
//
// Inboxes are double-buffered by epoch parity: a rank can be at most one exchange ahead of the slowest one (it
// cannot pass step 3 of exchange e+1 before everybody has sent e+1, i.e. finished reading e), so the buffer of
// parity e is never overwritten while somebody still reads it.  Epochs only grow, flags are never reset.
// k_peer_allreduce is the same push/flag/sum primitive for the one-off reductions (construction totals, final
// log-likelihood, reassignment column sums).
#pragma once
#include "tsc_kernels.cuh"

namespace tsc {

constexpr int kTailThreads = 256;
constexpr int kTailLoci = 4;                              // loci per thread
static_assert(kTailLoci == 4, "k_tail initialises four sums");
constexpr int kTailBlockLoci = kTailThreads * kTailLoci;  // 1024 loci per block

struct PeerArgs {
    unsigned char* const* bufs;   // device array: the `world` exchange buffers as mapped in this process
    int world, rank;              // rank = this GPU's rank in [0, world)
    int kpad, nb_max, cap;        // K rounded up to 32; flag slots per rank; 8-byte words per generic inbox
};

__host__ __device__ inline size_t peer_buffer_words(int world, int kpad, int nb_max, int cap) {
    return (size_t)world * nb_max + world + 2ULL * world * kpad + 2ULL * world * cap;
}
__device__ __forceinline__ unsigned long long* peer_flags_iter(unsigned char* buf) {
    return reinterpret_cast<unsigned long long*>(buf);
}
__device__ __forceinline__ unsigned long long* peer_flags_gen(const PeerArgs& p, unsigned char* buf) {
    return peer_flags_iter(buf) + (size_t)p.world * p.nb_max;
}
__device__ __forceinline__ double* peer_inbox_iter(const PeerArgs& p, unsigned char* buf, int parity, int r) {
    return reinterpret_cast<double*>(peer_flags_gen(p, buf) + p.world) + ((size_t)parity * p.world + r) * p.kpad;
}
__device__ __forceinline__ unsigned long long* peer_inbox_gen(const PeerArgs& p, unsigned char* buf, int parity, int r) {
    return reinterpret_cast<unsigned long long*>(peer_inbox_iter(p, buf, 0, 0) + 2ULL * p.world * p.kpad) +
           ((size_t)parity * p.world + r) * p.cap;
}

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_gpu(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= epoch; false after ~20 s (a peer died): the caller records the error instead of hanging the GPU
template <bool GPU_SCOPE>
__device__ __forceinline__ bool peer_wait(const unsigned long long* flag, unsigned long long epoch) {
    auto poll = [&]() { return GPU_SCOPE ? ld_acquire_gpu(flag) : ld_acquire_sys(flag); };
    if (poll() >= epoch) return true;
    const unsigned long long t0 = globaltimer_ns();
    for (;;) {
        for (int i = 0; i < 64; ++i)
            if (poll() >= epoch) return true;
        if (globaltimer_ns() - t0 > 20000000000ULL) return false;
        __nanosleep(200);
    }
}

struct TailArgs {
    UpdateArgs u;
    PeerArgs p;
    double* acc;                    // R replicas of K doubles
    int R;
    unsigned long long epoch_base;  // the exchange of iteration i (0-based, device-side counter) has epoch base + i + 1,
                                    // so ranks agree on it however far ahead each host has queued launches
    double* partials;               // one |pi' - pi| partial per block
    unsigned* ticket;               // blocks finished (reset by the last one)
};

__global__ void __launch_bounds__(kTailThreads) k_tail(const TailArgs a) {
    __shared__ double s_red[32];
    __shared__ int s_last;
    EmState* st = a.u.st;
    if (st->done) return;                      // uniform: only the last block of the previous launch writes it
    const int K = a.u.K, tid = threadIdx.x;
    const int iter = st->iter;
    const unsigned long long epoch = a.epoch_base + (unsigned long long)iter + 1ULL;
    const int parity = (int)(epoch & 1ULL);
    unsigned char* own = a.p.bufs[a.p.rank];
    // ---- 1 + 2: this GPU's sums of the block's loci, pushed to every rank
    double s[kTailLoci] = {0.0, 0.0, 0.0, 0.0};
    {   // all loads of a replica row are independent: keep 4 loci x 4 replicas in flight per thread
        const int j0 = blockIdx.x * kTailBlockLoci + tid;
#pragma unroll 4
        for (int r = 0; r < a.R; ++r) {
            double* row = a.acc + (size_t)r * K;
#pragma unroll
            for (int i = 0; i < kTailLoci; ++i) {
                const int j = j0 + i * kTailThreads;
                if (j < K) { s[i] += row[j]; row[j] = 0.0; }
            }
        }
    }
    for (int q = 0; q < a.p.world; ++q) {
        double* in = peer_inbox_iter(a.p, a.p.bufs[q], parity, a.p.rank);
#pragma unroll
        for (int i = 0; i < kTailLoci; ++i) {
            const int j = blockIdx.x * kTailBlockLoci + i * kTailThreads + tid;
            if (j < K) in[j] = s[i];
        }
    }
    // the block's stores happen-before the barrier, the flag's release (cumulative) publishes them: system scope when
    // other GPUs read them, device scope (several microseconds cheaper) when the model lives on one GPU
    __syncthreads();
    const bool solo = a.p.world == 1;
    if (tid < a.p.world) {
        unsigned long long* f = peer_flags_iter(a.p.bufs[tid]) + (size_t)a.p.rank * a.p.nb_max + blockIdx.x;
        if (solo) st_release_gpu(f, epoch); else st_release_sys(f, epoch);
    }
    // ---- 3: every rank's every block has delivered (identical loci of another block may be needed: `rep`)
    bool ok = true;
    for (int f = tid; f < a.p.world * (int)gridDim.x; f += kTailThreads) {
        const int r = f / (int)gridDim.x, b = f - r * (int)gridDim.x;
        const unsigned long long* fl = peer_flags_iter(own) + (size_t)r * a.p.nb_max + b;
        ok = (solo ? peer_wait<true>(fl, epoch) : peer_wait<false>(fl, epoch)) && ok;
    }
    if (!ok) st->pad = 1;                      // error flag, read by the host after the loop
    __syncthreads();
    // ---- 4 + 5
    const Consts c = *a.u.c;
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < kTailLoci; ++i) {
        const int j = blockIdx.x * kTailBlockLoci + i * kTailThreads + tid;
        if (j < K) {
            const int jj = a.u.rep ? a.u.rep[j] : j;            // identical loci share one sum -> exact ties survive
            double ts = 0.0;
            for (int r = 0; r < a.p.world; ++r) ts += __ldcg(peer_inbox_iter(a.p, own, parity, r) + jj);
            const double th = (ts + c.theta_prior_wt) / c.theta_denom;
            const double pisum = a.u.pisum0[j] + ts;
            const double pn = (pisum + c.pi_prior_wt) / c.pi_denom;
            const double po = a.u.pi[j];
            local += fabs(pn - po);
            a.u.pi_prev[j] = po; a.u.theta_prev[j] = a.u.theta[j]; a.u.pt_prev[j] = a.u.pt[j];
            a.u.pi[j] = pn; a.u.theta[j] = th; a.u.pt[j] = pn * th;
            if (iter == 0) { a.u.pi_init[j] = pn; a.u.theta_init[j] = th; }
        }
    }
    const double part = block_sum(local, s_red);
    if (tid == 0) {
        a.partials[blockIdx.x] = part;
        __threadfence();
        s_last = (atomicAdd(a.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- last block: diff_est in block order, loop control
    __threadfence();
    double d = 0.0;
    for (int b = tid; b < (int)gridDim.x; b += kTailThreads) d += __ldcg(a.partials + b);
    const double diff = block_sum(d, s_red);
    if (tid == 0) {
        *a.ticket = 0u;
        a.u.diffs[iter] = diff;
        st->diff = diff;
        st->iter = iter + 1;
        if (!a.u.use_lnl) {
            const int conv = diff < a.u.eps;
            st->converged = conv;
            st->done = conv || (iter + 1 >= a.u.max_iter);
        }
    }
}

// data[0..n) <- reduction over all ranks (rank order) of data[0..n); 8-byte elements, n <= cap.
// OP 0: f64 sum, 1: f64 max, 2: u64 sum.  One block.
template <int OP>
__global__ void __launch_bounds__(1024) k_peer_allreduce(const PeerArgs p, unsigned long long* __restrict__ data, int n,
                                                          unsigned long long epoch, int* __restrict__ err) {
    const int parity = (int)(epoch & 1ULL), tid = threadIdx.x;
    unsigned char* own = p.bufs[p.rank];
    for (int q = 0; q < p.world; ++q) {
        unsigned long long* in = peer_inbox_gen(p, p.bufs[q], parity, p.rank);
        for (int i = tid; i < n; i += 1024) in[i] = data[i];
    }
    __syncthreads();
    if (tid < p.world) st_release_sys(peer_flags_gen(p, p.bufs[tid]) + p.rank, epoch);
    if (tid < p.world && !peer_wait<false>(peer_flags_gen(p, own) + tid, epoch)) *err = 1;
    __syncthreads();
    for (int i = tid; i < n; i += 1024) {
        unsigned long long acc = __ldcg(peer_inbox_gen(p, own, parity, 0) + i);
        for (int r = 1; r < p.world; ++r) {
            const unsigned long long v = __ldcg(peer_inbox_gen(p, own, parity, r) + i);
            if (OP == 0) acc = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v));
            else if (OP == 1) acc = (unsigned long long)__double_as_longlong(fmax(__longlong_as_double((long long)acc), __longlong_as_double((long long)v)));
            else acc += v;
        }
        data[i] = acc;
    }
}

}  // namespace tsc
