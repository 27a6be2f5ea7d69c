/*
 * telescope_b200.h -- C ABI of libtelescope_b200.so: Telescope's EM reassignment loop on NVIDIA B200 (sm_100a).
 *
 * The reference (mlbendall/telescope @ 4cf18595) has no FFI: its seam for this path is the Python class
 * `TelescopeLikelihood` (telescope/utils/model.py:631-865) on top of `csr_matrix_plus`
 * (telescope/utils/sparse_plus.py:24-174).  Each entry point below names the reference member it replaces; the
 * ctypes binding a maintainer adds on the reference side is shown in INTEGRATION.md and shipped as
 * telescope_b200/_abi.py + telescope_b200/likelihood.py (class TelescopeLikelihood, same constructor, methods and
 * attributes).
 *
 * Conventions
 *   - plain pointers and sizes only; the caller owns every host buffer, the handle owns all device memory,
 *     streams, events and NCCL communicators;
 *   - every call returns 0 on success or a non-zero tsc_status; tsc_last_error() gives the message
 *     (thread-local).  There is NO CPU fallback: without a usable CUDA device every compute call fails;
 *   - a handle is driven by one host thread; calls block until their results are in the caller's buffers;
 *   - rows are reads, columns are loci.  The score matrix is CSR: indptr[n_rows+1], indices[nnz] (int32,
 *     0 <= j < n_cols), raw[nnz] (uint16 alignment scores).  Per-entry outputs (z, assignment data) come back in
 *     exactly that entry order, with an explicit 0.0 where the reference's scipy result would not store an entry.
 *
 * Sharding: reads are split into contiguous row ranges, one shard per GPU, balanced by nnz.  One handle can drive
 * several GPUs of the box from one process (n_local_devices > 1), and/or be one of n_procs cooperating processes
 * (one per GPU under torchrun); either way the only data-path exchange is the K per-locus M-step sums, once per EM
 * iteration (plus a few one-off reductions at construction and after the loop) -- through peer-mapped memory inside
 * the update kernel on one node (tsc_transport), or an NCCL all-reduce.
 */
#ifndef TELESCOPE_B200_H
#define TELESCOPE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TSC_ABI_VERSION 2

typedef struct tsc_handle tsc_handle;

typedef enum {
    TSC_OK = 0,
    TSC_ERR_ARG = 1,      /* bad argument (also: unknown reassign method -> Python raises ValueError, model.py:834-835) */
    TSC_ERR_CUDA = 2,     /* CUDA runtime error, including "no device" */
    TSC_ERR_NCCL = 3,     /* NCCL missing or failed */
    TSC_ERR_STATE = 4,    /* call out of order (e.g. results requested before tsc_em) */
    TSC_ERR_ALLOC = 5
} tsc_status;

/* reassign methods, model.py:834 */
typedef enum {
    TSC_EXCLUDE = 0, TSC_CHOOSE = 1, TSC_AVERAGE = 2, TSC_CONF = 3, TSC_UNIQUE = 4, TSC_ALL = 5
} tsc_method;

/* fused E+M kernel variants (tsc_config.kernel) */
typedef enum {
    TSC_KERNEL_AUTO = 0,
    TSC_KERNEL_ROWS = 1,  /* one sub-warp per read; simple, used as the in-library cross-check */
    TSC_KERNEL_TILES = 2, /* flat 128-entry tiles with warp segmented scan (round-1 fast path; still used for the
                             posterior / log-likelihood passes and for reads that do not fit a slice) */
    TSC_KERNEL_ELL = 3    /* locus-clustered sliced-ELL stream, one read per lane, per-warp locus window in shared
                             memory, TMA-fed (AUTO selects this) */
} tsc_kernel;

typedef struct {
    int32_t n_local_devices;   /* GPUs this process drives (>= 1) */
    const int32_t* device_ids; /* CUDA ordinals, n_local_devices of them; NULL = 0..n-1 */
    int32_t n_procs;           /* cooperating processes (1 = single process) */
    int32_t proc_rank;         /* this process's rank in [0, n_procs) */
    const void* nccl_id;       /* 128-byte id from tsc_nccl_unique_id() of rank 0; required iff n_procs > 1 */
    int32_t kernel;            /* tsc_kernel */
    int32_t replicas;          /* copies of the per-locus accumulator in L2 (0 = auto) */
    int32_t smem_table_cols;   /* loci whose pi*theta entry is staged in shared memory (-1 = auto = 0: the L1/L2
                                  gather measured faster, profiles/r1_microbench_scatter_gather.log) */
    int32_t smem_acc_cols;     /* reserved (shared-memory fp64 atomics are CAS loops on sm_100a; not used) */
    int32_t permute_columns;   /* 1 = renumber loci by descending entry count internally, 0 = keep the caller's
                                  numbering (default: neighbouring loci share sectors, which the scatter-add likes) */
    int32_t transport;         /* tsc_transport: how the GPUs exchange the K per-locus sums of an iteration */
    void* peer_buffer;         /* peer transport, n_procs > 1: this process's exchange buffer from
                                  tsc_peer_buffer_create (the handle takes ownership) */
    const void* peer_handles;  /* ... and the 64-byte CUDA IPC handles of all n_procs buffers, in rank order */
    int32_t reserved[4];
} tsc_config;

/* exchange of the per-locus M-step sums between the GPUs of a model */
typedef enum {
    TSC_TRANSPORT_AUTO = 0,   /* peer when every GPU can map every other one's buffer, else NCCL */
    TSC_TRANSPORT_PEER = 1,   /* one node: peer-mapped exchange buffers (peer access inside one process, CUDA IPC
                                 between processes); the exchange is fused into the update kernel, no NCCL involved */
    TSC_TRANSPORT_NCCL = 2    /* ncclAllReduce over a communicator built from nccl_id */
} tsc_transport;

/* library / environment */
int tsc_abi_version(void);
const char* tsc_last_error(void);
int tsc_device_count(int32_t* n_out);
/* path of libnccl.so.2 to dlopen (optional; default: $TELESCOPE_B200_NCCL, else the loader's search path) */
int tsc_set_nccl_path(const char* path);
int tsc_nccl_unique_id(void* out128);
/* Peer transport with one process per GPU: allocate this rank's exchange buffer on `device` for a model with n_cols
 * loci and `world` ranks, and export its 64-byte CUDA IPC handle.  The ranks swap the handles (any side channel; a
 * file all-gather is in telescope_b200/dist.py) and pass buffer + all handles to tsc_create through tsc_config. */
int tsc_peer_buffer_create(int32_t device, int32_t n_cols, int32_t world, void** buf_out, void* ipc_handle64_out);
void tsc_peer_buffer_free(void* buf);     /* only for a buffer that never reached tsc_create */
/* transport the handle ended up with (tsc_transport) */
int tsc_get_transport(tsc_handle* h, int32_t* transport_out);
void tsc_config_default(tsc_config* cfg);
/* page-locked host memory for the CSR arrays (optional; uploads from it run at full PCIe speed) */
void* tsc_pinned_alloc(uint64_t bytes);
void tsc_pinned_free(void* p);
/* Device blocks of destroyed models are kept for the next model of this process (cudaMalloc / cudaFree of tens of
 * gigabytes cost more than the upload they precede); this hands them back to the driver.  The cache is bounded by
 * TELESCOPE_B200_CACHE_GB (default 96, 0 = off) and emptied on its own when an allocation fails.  No counterpart in the
 * reference (numpy frees on garbage collection). */
void tsc_trim_memory(void);
/* Host helper of reassign('choose') / Telescope.output_report's init_best_random: n bounded draws of numpy's legacy
 * global generator, element i uniform in [0, counts[i]), exactly as a sequence of np.random.choice(range(a, b)) calls
 * consumes it (sparse_plus.py:146-153).  key624 / pos: np.random.get_state()[1:3], updated in place for set_state().
 * counts[i] <= 1 gives pick 0 and consumes nothing; counts must be below 2^31.  No device involved. */
int tsc_mt19937_draw_picks(uint32_t* key624, int32_t* pos, const int64_t* counts, int64_t n, int32_t* picks);
/* the same over a per-read array of best-hit counts (tsc_reassign_nbest / tsc_report): reads with at most one best hit get
 * pick 0 and consume nothing, so the stream equals the reference's loop over the tie reads only */
int tsc_mt19937_draw_rows(uint32_t* key624, int32_t* pos, const int32_t* nbest, int64_t n, int32_t* picks);

/*
 * TelescopeLikelihood.__init__ (model.py:635-700).
 *   q_lut[s], s in [0, lut_len): the host-computed table expm1((s * (1/max_score)) * 100.) (model.py:652-653),
 *   computed with numpy so that Q is bit-identical to the reference's; every raw score must be < lut_len.
 *   indptr_bytes is 4 or 8 (scipy hands out int32 below 2^31 entries, int64 above).
 * With n_procs > 1 every process passes its own contiguous block of reads (local indptr starting at 0), the same
 * n_cols, q_lut and priors; max_score must already be the global maximum.
 * Builds Q, Y, the read weights w = max_j Q_ij, total/ambiguous weight, the weighted priors and pisum0 on the
 * device(s) and sets pi = theta = 1/K.
 */
int tsc_create(tsc_handle** out, const tsc_config* cfg,
               int64_t n_rows, int32_t n_cols, int64_t nnz,
               const void* indptr, int32_t indptr_bytes, const int32_t* indices, const uint16_t* raw,
               const double* q_lut, int32_t lut_len, double pi_prior, double theta_prior);
void tsc_destroy(tsc_handle* h);

/* scalars of __init__ (model.py:690-697): total_wt, ambig_wt, max weight, pi_prior_wt, theta_prior_wt; pisum0[K] */
int tsc_get_constants(tsc_handle* h, double* scalars5, double* pisum0);
/* Y (model.py:679) and w (model.py:690) per local read, either pointer may be NULL */
int tsc_get_row_info(tsc_handle* h, uint8_t* y_rows, double* w_rows);
/* Q data (model.py:653) in entry order */
int tsc_get_q(tsc_handle* h, double* q_data);

/*
 * TelescopeLikelihood.em (model.py:762-806): runs estep/mstep until |pi' - pi|_1 < eps (or |dlnl| < eps with
 * use_likelihood) or max_iter, entirely on the device(s).  diffs_out / lnls_out (max_iter doubles each, lnls_out
 * may be NULL unless use_likelihood) receive the per-iteration values the reference logs (model.py:787,791).
 * Afterwards the handle holds pi, theta, pi_init, theta_init, the parameters of the last E-step (so that the
 * posterior z of model.py:795 can be regenerated on demand) and lnl (model.py:800-801).
 */
int tsc_em(tsc_handle* h, int32_t max_iter, double eps, int32_t use_likelihood,
           double* diffs_out, double* lnls_out, int32_t* n_iter, int32_t* converged, double* final_lnl);
/* per-iteration device time (ms, CUDA events) of the fused E+M kernel of the last tsc_em on local shard 0 */
int tsc_get_kernel_times(tsc_handle* h, float* ms_out, int32_t max_n, int32_t* n_out);
/* ... and of what follows it in the iteration: replica sum, exchange between the GPUs, parameter update, loop control
 * (one kernel with the peer transport; reduce + ncclAllReduce + update with NCCL) */
int tsc_get_tail_times(tsc_handle* h, float* ms_out, int32_t max_n, int32_t* n_out);
/* device time (ms, CUDA events on the library's stream of local shard 0) of the whole last tsc_em loop, first
 * kernel to the final log-likelihood reduction */
int tsc_get_em_device_ms(tsc_handle* h, float* ms_out);
/*
 * Diagnostic: run one device pass `reps` times on local shard 0 without any host copy and return the mean device
 * time (CUDA events).  pass_id: 0 = fused E+M kernel (accumulators are discarded), 1 = E-step alone writing z
 * (model.py:702-722), 2 = log-likelihood pass (model.py:744-760), 3 = reassign column sums, mode exclude.
 * Does not change pi/theta or the EM state.
 */
int tsc_time_pass(tsc_handle* h, int32_t pass_id, int32_t reps, float* mean_ms);
/* Diagnostic: host-side wall time of the stages of tsc_create, "stage=ms;stage=ms;..." (valid until tsc_destroy) */
const char* tsc_create_laps(tsc_handle* h);
/* All-reduce n (<= 8) host doubles over every shard of every process of this handle's communicator (op 0 = sum,
 * 1 = max); with n_procs == 1 and one device it is the identity.  Lets multi-process drivers agree on timings and
 * synchronise without a second communication library. */
int tsc_allreduce_f64(tsc_handle* h, double* inout, int32_t n, int32_t op);
/* Diagnostic: device layout of local shard 0 as 8 int64: [0] bytes of the clustered slice stream of the fused kernel,
 * [1] slices, [2] reads in the stream, [3] stored entries in the stream, [4] reads / [5] entries of the residual CSR
 * (ambiguous reads that do not fit the stream), [6] CTAs of the stream kernel, [7] long-read records (more than 48 entries) */
int tsc_get_layout_stats(tsc_handle* h, int64_t* out8);
/* launches = kernels this library launched since creation; bytes moved over PCIe by this handle */
int tsc_get_counters(tsc_handle* h, int64_t* launches, int64_t* h2d_bytes, int64_t* d2h_bytes);

/* pi, theta, pi_init, theta_init (model.py:777-778,796); any pointer may be NULL */
int tsc_get_params(tsc_handle* h, double* pi, double* theta, double* pi_init, double* theta_init);
/* overwrite the current pi/theta (resets nothing else); used to drive single steps from tests */
int tsc_set_params(tsc_handle* h, const double* pi, const double* theta);

/* TelescopeLikelihood.estep(pi, theta) (model.py:702-722): z for the local entries */
int tsc_estep(tsc_handle* h, const double* pi, const double* theta, double* z_data);
/* TelescopeLikelihood.mstep(z) (model.py:724-742): z_data for the local entries -> pi_hat, theta_hat (global) */
int tsc_mstep(tsc_handle* h, const double* z_data, double* pi_hat, double* theta_hat);
/* TelescopeLikelihood.calculate_lnl(z, pi, theta) (model.py:744-760), global */
int tsc_calculate_lnl(tsc_handle* h, const double* z_data, const double* pi, const double* theta, double* lnl);

/* self.z after em() (model.py:795), or Q.norm(1) when initial != 0 (model.py:837) */
int tsc_get_z(tsc_handle* h, int32_t initial, double* z_data);

/*
 * TelescopeLikelihood.reassign (model.py:808-865).
 * tsc_reassign_nbest: number of best hits per local read (sparse_plus.py:99-129); reads with more than one consume
 *   one draw of the caller's RNG for TSC_CHOOSE (sparse_plus.py:146-153) -- the host draws, in read order.
 * picks: for TSC_CHOOSE, picks[i] in [0, nbest_i) selects which best hit read i keeps; ignored otherwise (NULL ok).
 * tsc_reassign_colsum: reassign(...).sum(0) as doubles (integers are exact), global over all shards.
 * tsc_reassign_data: the assignment matrix's data for the local entries (0.0 where nothing is stored).
 */
int tsc_reassign_nbest(tsc_handle* h, int32_t initial, int32_t* nbest_rows);
/*
 * Telescope.output_report (model.py:432-458) in one pass over the matrix instead of seven reassign() calls.
 * out6k = 6 vectors of n_cols doubles, global over all shards:
 *   [0] reassign('conf', thresh)                     final_conf
 *   [1] reassign('all', initial=True)                init_aligned
 *   [2] reassign('unique')                           unique_count
 *   [3] reassign('exclude', initial=True)            init_best
 *   [4] reassign('average', initial=True)            init_best_avg
 *   [5] reassign(final_method, thresh)               the counts file; for TSC_CHOOSE only the reads with one best hit
 * nbest_init / nbest_final (optional, per local read): best hits of the initial / final posterior, for the host's
 * RNG draws of 'choose'; tsc_choose_ties_colsum then returns the column sums of the drawn hits of the tie reads
 * (reads with nbest > 1), to be added to [3] (init_best_random) or [5].
 */
int tsc_report(tsc_handle* h, double thresh, int32_t final_method, int32_t* nbest_init, int32_t* nbest_final,
               double* out6k);
int tsc_choose_ties_colsum(tsc_handle* h, int32_t initial, const int32_t* nbest, const int32_t* picks, double* colsum);
int tsc_reassign_colsum(tsc_handle* h, int32_t method, double thresh, int32_t initial,
                        const int32_t* picks, double* colsum);
int tsc_reassign_data(tsc_handle* h, int32_t method, double thresh, int32_t initial,
                      const int32_t* picks, double* data);

#ifdef __cplusplus
}
#endif
#endif /* TELESCOPE_B200_H */
