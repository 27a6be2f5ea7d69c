// libtelescope_b200.so -- host orchestration and C ABI (include/telescope_b200.h).
//
// Replaces, behind a C ABI, TelescopeLikelihood (reference telescope/utils/model.py:631-865) and the
// csr_matrix_plus helpers it calls (telescope/utils/sparse_plus.py:16-165).  No CPU fallback: every compute entry
// point needs a CUDA device and fails loudly otherwise.
//
// Layout in HBM, per shard (= one GPU's contiguous block of reads):
//   q[nnz] fp64, col[nnz] int32 (the caller's locus numbering), indptr[rows+1] int64, wy[rows] fp64 (= w_i * Y_i),
//   tiles[~nnz/115] 32-byte descriptors of the flat-tile passes (with the clustered stream: built when the posterior
//   is first exported); the clustered slice stream + its record index (what the per-iteration kernel reads) and the
//   residual CSR of the reads outside it; K-length fp64 vectors: pi, theta,
//   pt (= pi*theta), their *_prev twins (the parameters the stored posterior z was computed from, model.py:795),
//   *_init, pisum0, accumulator replicas; the exchange buffer every rank of the node maps.
// Per EM iteration and shard: stream kernel (+ flat tiles on the residual front) -> k_tail (replica sum, exchange
// between the GPUs through peer memory, update, loop control).  The loop runs ahead of the host; convergence is decided
// on the device and polled through pinned memory.  The first iterations of a model also time every CTA of the stream
// kernel and re-partition the stream by the measured time (k_ell_rebalance).
// Construction: the entry arrays are uploaded in chunks and a second stream runs each chunk's kernels beside the next
// chunk's copy; device blocks, exchange buffers and IPC mappings of destroyed models are kept for the next model of the
// process (DevCache, IpcCache).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <functional>
#include <limits>
#include <mutex>
#include <numeric>
#include <string>
#include <thread>
#include <map>
#include <unordered_map>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>   // types and prototypes only; the library is dlopen'ed (tsc_set_nccl_path)

#include "../../include/telescope_b200.h"
#include "tsc_kernels.cuh"
#include "tsc_tiles.cuh"
#include "tsc_ell.cuh"
#include "tsc_peer.cuh"

using namespace tsc;

// ------------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }

#define CU(x)                                                                                         \
    do {                                                                                              \
        cudaError_t e_ = (x);                                                                         \
        if (e_ != cudaSuccess)                                                                        \
            return fail(TSC_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                          std::to_string(__LINE__) + ")");                            \
    } while (0)

// ------------------------------------------------------------------------------------------------- NCCL (dlopen)
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;
static std::string g_nccl_path;

static int nccl_load() {
    if (g_nccl.lib) return TSC_OK;
    std::vector<std::string> cands;
    if (!g_nccl_path.empty()) cands.push_back(g_nccl_path);
    if (const char* e = getenv("TELESCOPE_B200_NCCL")) cands.push_back(e);
    cands.push_back("libnccl.so.2");
    cands.push_back("libnccl.so");
    std::string tried;
    for (auto& c : cands) {
        void* lib = dlopen(c.c_str(), RTLD_NOW | RTLD_GLOBAL);
        if (!lib) { tried += c + " (" + (dlerror() ? "not loadable" : "?") + "); "; continue; }
        NcclApi a; a.lib = lib;
        a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
        a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
        a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
        a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
        a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
        a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllReduce && a.GroupStart && a.GroupEnd && a.GetErrorString) {
            g_nccl = a;
            return TSC_OK;
        }
        tried += c + " (symbols missing); ";
        dlclose(lib);
    }
    return fail(TSC_ERR_NCCL, "cannot load NCCL: " + tried);
}

#define NC(x)                                                                                         \
    do {                                                                                              \
        ncclResult_t r_ = (x);                                                                        \
        if (r_ != ncclSuccess)                                                                        \
            return fail(TSC_ERR_NCCL, std::string(#x) + ": " + g_nccl.GetErrorString(r_) + " (" + __FILE__ + ":" + \
                                          std::to_string(__LINE__) + ")");                            \
    } while (0)

// ------------------------------------------------------------------------------------------------- device memory
// cudaMalloc / cudaFree of large blocks cost milliseconds each -- sometimes tens of milliseconds while copies are in
// flight -- and cudaFree synchronises the device.  A model allocates ~15 blocks; a process that builds models one after
// the other (the CLI's assign + resume, a service, the benchmark's repeated jobs) would pay that every time.  So blocks
// go back to a process-wide cache when a model is destroyed and are handed out again to a request of
// (nearly) the same size on the same device (model after model on similar data asks for the same sizes).  The cache is
// bounded (TELESCOPE_B200_CACHE_GB, default 96; 0 switches it off), emptied by tsc_trim_memory() and whenever an
// allocation fails.  Pinned host words of the loop state come from a small pool that is never returned.
struct DevCache {
    std::mutex mu;
    std::unordered_map<void*, std::pair<int, size_t>> live;                 // handed out: device, bytes
    std::multimap<std::pair<int, size_t>, void*> idle;                      // (device, bytes) -> block
    size_t idle_bytes = 0;
    std::vector<void*> pinned_words;                                        // 256-byte pinned host slots, reusable
    double ms = 0.0;                                                        // time spent in the driver's allocation calls
    long long calls = 0, hits = 0;
    size_t limit() const {
        const char* e = getenv("TELESCOPE_B200_CACHE_GB");
        const double gb = e ? atof(e) : 96.0;
        return gb <= 0 ? 0 : (size_t)(gb * (double)(1ULL << 30));
    }
    void trim_locked(size_t keep) {
        while (idle_bytes > keep && !idle.empty()) {
            auto it = std::prev(idle.end());                                // largest block of the highest device first
            int cur = 0;
            cudaGetDevice(&cur);
            cudaSetDevice(it->first.first);
            cudaFree(it->second);
            cudaSetDevice(cur);
            idle_bytes -= it->first.second;
            idle.erase(it);
        }
    }
};
static DevCache g_cache;
constexpr size_t kCacheMinBytes = 1;          // every block: even a 4-byte cudaMalloc / cudaFree pair is a device synchronisation

static cudaError_t dev_malloc_bytes(void** out, size_t bytes) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (bytes >= kCacheMinBytes) {
        std::lock_guard<std::mutex> g(g_cache.mu);
        // best fit: the smallest idle block of this device that is large enough and at most 1/8 (or 64 KB) larger -- a few
        // sizes of a model depend on the order in which atomics happened to append reads and differ by a little
        auto it = g_cache.idle.lower_bound({dev, bytes});
        if (it != g_cache.idle.end() && it->first.first == dev && it->first.second <= bytes + std::max<size_t>(bytes / 8, 65536)) {
            const size_t got = it->first.second;
            *out = it->second;
            g_cache.idle_bytes -= got;
            g_cache.idle.erase(it);
            g_cache.live[*out] = {dev, got};
            ++g_cache.hits;
            return cudaSuccess;
        }
    }
    const auto t0 = std::chrono::steady_clock::now();
    cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaErrorMemoryAllocation) {           // make room: everything idle goes back to the driver, once
        cudaGetLastError();
        { std::lock_guard<std::mutex> g(g_cache.mu); g_cache.trim_locked(0); }
        e = cudaMalloc(out, bytes);
    }
    std::lock_guard<std::mutex> g(g_cache.mu);
    g_cache.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    ++g_cache.calls;
    if (e == cudaSuccess && bytes >= kCacheMinBytes) g_cache.live[*out] = {dev, bytes};
    return e;
}
template <typename T> static cudaError_t dev_malloc(T** out, size_t bytes) { return dev_malloc_bytes((void**)out, bytes); }

// Like cudaFree, this waits for the device first: nobody may still be using the block when its next owner gets it.
static void dev_free(void* p) {
    if (!p) return;
    std::unique_lock<std::mutex> g(g_cache.mu);
    auto it = g_cache.live.find(p);
    if (it == g_cache.live.end()) {
        g.unlock();
        const auto t0 = std::chrono::steady_clock::now();
        cudaFree(p);
        g.lock();
        g_cache.ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        ++g_cache.calls;
        return;
    }
    const std::pair<int, size_t> key = it->second;
    g_cache.live.erase(it);
    const size_t lim = g_cache.limit();
    if (lim == 0 || key.second > lim) { g.unlock(); cudaFree(p); return; }
    g.unlock();
    {   // the block's last user may still be running
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != key.first) cudaSetDevice(key.first);
        cudaDeviceSynchronize();
        if (cur != key.first) cudaSetDevice(cur);
    }
    g.lock();
    g_cache.idle.insert({key, p});
    g_cache.idle_bytes += key.second;
    if (g_cache.idle_bytes > lim) g_cache.trim_locked(lim);
}

static void* pinned_word_take() {
    {
        std::lock_guard<std::mutex> g(g_cache.mu);
        if (!g_cache.pinned_words.empty()) { void* p = g_cache.pinned_words.back(); g_cache.pinned_words.pop_back(); return p; }
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, 256, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
static void pinned_word_give(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(g_cache.mu);
    g_cache.pinned_words.push_back(p);
}

extern "C" void tsc_trim_memory(void) {
    std::lock_guard<std::mutex> g(g_cache.mu);
    g_cache.trim_locked(0);
}

// ------------------------------------------------------------------------------------------------- handle
struct Shard {
    int dev = 0;
    int world_rank = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux = nullptr;        // construction: the per-chunk kernels that run beside the upload of the next chunk
    cudaEvent_t ev_up = nullptr;
    long long n_rows = 0, nnz = 0;
    long long row_begin = 0, nnz_begin = 0;   // within this process's (compacted) reads / entries
    long long* indptr = nullptr;
    int* col = nullptr;
    double* q = nullptr;
    double* wy = nullptr;
    Tile* tiles = nullptr;
    long long n_tiles = 0, n_long = 0;
    // K-vectors
    double *pi = nullptr, *theta = nullptr, *pt = nullptr, *pi_prev = nullptr, *theta_prev = nullptr, *pt_prev = nullptr;
    double *pi_init = nullptr, *theta_init = nullptr, *pisum0 = nullptr, *acc = nullptr, *thetasum = nullptr;
    double *ones = nullptr, *tmp_a = nullptr, *tmp_b = nullptr, *tmp_c = nullptr, *colsum = nullptr;
    int* perm = nullptr;      // original -> internal locus index (nullptr = identity)
    int* rep = nullptr;       // internal locus -> representative of its class of identical columns (nullptr = none)
    Consts* consts = nullptr;
    EmState* st = nullptr;
    EmState* st_host = nullptr;    // pinned, 2 slots
    double* diffs = nullptr;
    double* lnls = nullptr;
    int diffs_cap = 0;
    double* partials = nullptr;    // per-block partial sums (lnl)
    double* scalars = nullptr;     // [0..3] init partials / lnl
    int* bad = nullptr;
    ncclComm_t comm = nullptr;
    cudaEvent_t ev_poll[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_k;   // 3 per timed iteration, shard 0 only: before / after the fused kernels, after the tail
    cudaEvent_t ev_em[2] = {nullptr, nullptr};
    int n_sm = 0;
    int grid_rows = 0, grid_tiles = 0;
    size_t smem_tiles = 0;
    int s_cols = 0;
    // device memory comes in a few slabs (one cudaMalloc each: allocation calls cost milliseconds and their cost
    // varies wildly); the array pointers of this struct point into them unless allocated on their own
    std::vector<std::pair<char*, size_t>> slabs;
    bool in_slab(const void* p) const {
        for (auto& b : slabs) if ((const char*)p >= b.first && (const char*)p < b.first + b.second) return true;
        return false;
    }
    uint16_t* raw = nullptr;                    // uploaded scores; after construction: scratch of the clustering
    // peer-memory transport (tsc_peer.cuh): this GPU's exchange buffer, every rank's buffer as mapped here
    unsigned char* peer_own = nullptr;
    std::vector<unsigned char*> peer_map;       // world entries; [world_rank] == peer_own
    std::vector<char> peer_opened;              // mapped with cudaIpcOpenMemHandle by this model (kept open in g_ipc)
    unsigned char** peer_ptrs_d = nullptr;
    double* tail_partials = nullptr;
    unsigned* tail_ticket = nullptr;
    int* peer_err = nullptr;
    void* xchg_ptr = nullptr;                   // argument slot of the generic all-reduce for temporaries
    LogTab* log_tab = nullptr;                  // table of log1p_big (tsc_kernels.cuh)
    // clustered sliced-ELL stream of the fused kernel (tsc_ell.cuh) + residual CSR for the reads it does not hold
    unsigned char* ell_stream = nullptr;
    long long ell_bytes = 0, ell_slices = 0, ell_long = 0, ell_records = 0, ell_reads = 0, ell_entries = 0;
    int4* ell_index = nullptr;            // per slice record: offset / 16, first locus, T | last locus << 8, reads
    int ell_grid = 0, ell_grid_lnl = 0;
    long long *ell_range = nullptr, *ell_range_lnl = nullptr;      // record boundaries of the CTAs (grid + 1 each)
    unsigned* ell_cta_ns = nullptr;       // per CTA of the per-iteration kernel: duration of its last measured launch
    double* ell_rebal = nullptr;          // scratch of k_ell_rebalance, 2 * (grid + 1)
    int ell_rebal_left = 0;               // measured re-partitions still to do (the first iterations of the model)
    long long ell_slice_bytes = 0;        // bytes of the slice records (the long-read records follow them in the stream)
    long long *ell_lrange = nullptr, *ell_lrange_lnl = nullptr;    // the same for the long-read kernel
    int ell_lgrid = 0, ell_lgrid_lnl = 0;
    int* ell_rowid = nullptr;             // 16 per slice: the shard's read in each slot (-1 = empty)
    int* res_rowid = nullptr;             // per residual read: the shard's read
    long long res_amb_rows = 0, res_amb_nnz = 0;   // the ambiguous part of the residual CSR (it also holds the unique reads)
    long long res_rows = 0, res_nnz = 0, res_n_tiles = 0, res_n_long = 0;
    long long* res_indptr = nullptr;
    int* res_col = nullptr;
    double* res_q = nullptr;
    double* res_wy = nullptr;
    Tile* res_tiles = nullptr;            // over every read of the residual (log-likelihood)
    Tile* res_tiles_amb = nullptr;        // over its ambiguous front only (fused E+M)
    long long res_amb_tiles = 0, res_amb_long = 0;
};

struct tsc_handle {
    std::vector<Shard> shards;
    int K = 0, world = 1, n_procs = 1, proc_rank = 0;
    int R = 8, G = 8, kernel = TSC_KERNEL_TILES;     // R: accumulator replicas allocated
    int R_used = 8;                                  // ... and used (few when nearly everything goes through the stream)
    int transport = TSC_TRANSPORT_PEER;          // how the shards exchange K-vectors (TSC_TRANSPORT_*)
    int kpad = 0, nb_tail = 1, cap = 0;         // exchange-buffer geometry (tsc_peer.cuh)
    unsigned long long epoch_iter = 0, epoch_gen = 0;
    std::thread peer_thread;                    // maps the other ranks' exchange buffers while the CSR is uploaded
    int peer_thread_rc = TSC_OK;
    std::string peer_thread_err;
    bool smem_tab = false;
    long long n_rows_user = 0, n_rows = 0, nnz = 0;
    std::vector<long long> rowmap;        // compacted read -> caller's read index (empty when no empty reads)
    std::vector<int> perm, inv;           // perm[original] = internal; inv[internal] = original
    int n_dup_loci = 0;                   // loci whose column duplicates an earlier locus's
    std::string create_laps;              // host-side construction laps ("stage=ms;")
    std::vector<unsigned long long> pos_count;   // per locus (caller's numbering): stored entries with score > 0, global
    Consts consts{};
    std::vector<double> pisum0_host;
    bool em_done = false;
    int n_iter = 0, converged = 0;
    double lnl = std::numeric_limits<double>::infinity();
    std::vector<float> kernel_ms, tail_ms;
    float em_ms = 0.f;
    long long launches = 0, h2d = 0, d2h = 0;
};

#define LAUNCH(h) ((h)->launches++)

static inline int grid_for(long long work, int threads, int cap) {
    long long b = (work + threads - 1) / threads;
    return (int)std::max<long long>(1, std::min<long long>(b, cap));
}

static PeerArgs peer_args(const tsc_handle* h, const Shard& s) {
    return PeerArgs{s.peer_ptrs_d, h->world, s.world_rank, h->kpad, h->nb_tail, h->cap};
}

// the other ranks' exchange buffers are mapped (the mapping thread of setup_transport has finished)
static int peer_ready(tsc_handle* h) {
    if (h->peer_thread.joinable()) h->peer_thread.join();
    if (h->peer_thread_rc) return fail(h->peer_thread_rc, h->peer_thread_err);
    return TSC_OK;
}

// In-place reduction of `count` 8-byte elements over every shard of every process.
static int allreduce(tsc_handle* h, void* (*ptr_of)(Shard&), size_t count, ncclDataType_t dt, ncclRedOp_t op) {
    if (h->world == 1) return TSC_OK;
    { int rc = peer_ready(h); if (rc) return rc; }
    if (h->transport == TSC_TRANSPORT_NCCL) {
        NC(g_nccl.GroupStart());
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            void* p = ptr_of(s);
            NC(g_nccl.AllReduce(p, p, count, dt, op, s.comm, s.stream));
        }
        NC(g_nccl.GroupEnd());
        return TSC_OK;
    }
    const int kind = (dt == ncclUint64) ? 2 : (op == ncclMax ? 1 : 0);
    for (size_t done = 0; done < count;) {
        const int n = (int)std::min<size_t>(count - done, (size_t)h->cap);
        const unsigned long long epoch = ++h->epoch_gen;
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            unsigned long long* p = (unsigned long long*)ptr_of(s) + done;
            const PeerArgs pa = peer_args(h, s);
            if (kind == 0) k_peer_allreduce<0><<<1, 1024, 0, s.stream>>>(pa, p, n, epoch, s.peer_err);
            else if (kind == 1) k_peer_allreduce<1><<<1, 1024, 0, s.stream>>>(pa, p, n, epoch, s.peer_err);
            else k_peer_allreduce<2><<<1, 1024, 0, s.stream>>>(pa, p, n, epoch, s.peer_err);
            LAUNCH(h);
            CU(cudaGetLastError());
        }
        done += n;
    }
    return TSC_OK;
}
#define ALLREDUCE(h, member_expr, count, dt, op)                                                   \
    do {                                                                                           \
        int rc_ = allreduce((h), [](Shard& s) -> void* { return (void*)(member_expr); }, (count), (dt), (op)); \
        if (rc_) return rc_;                                                                       \
    } while (0)

static int sync_all(tsc_handle* h) {
    for (auto& s : h->shards) { CU(cudaSetDevice(s.dev)); CU(cudaStreamSynchronize(s.stream)); }
    return TSC_OK;
}

// K-vector in the caller's locus numbering -> device (internal numbering), and back
static int put_kvec(tsc_handle* h, Shard& s, const double* host, double* dev) {
    std::vector<double> tmp(h->K);
    for (int j = 0; j < h->K; ++j) tmp[h->perm[j]] = host[j];
    CU(cudaMemcpyAsync(dev, tmp.data(), sizeof(double) * h->K, cudaMemcpyHostToDevice, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    h->h2d += sizeof(double) * h->K;
    return TSC_OK;
}
static int get_kvec(tsc_handle* h, Shard& s, const double* dev, double* host) {
    std::vector<double> tmp(h->K);
    CU(cudaMemcpyAsync(tmp.data(), dev, sizeof(double) * h->K, cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    for (int j = 0; j < h->K; ++j) host[j] = tmp[h->perm[j]];
    h->d2h += sizeof(double) * h->K;
    return TSC_OK;
}

template <class F>
static int launch_rows(int G, F&& f) {
    switch (G) {
        case 4: f(std::integral_constant<int, 4>()); break;
        case 8: f(std::integral_constant<int, 8>()); break;
        case 16: f(std::integral_constant<int, 16>()); break;
        default: f(std::integral_constant<int, 32>()); break;
    }
    return TSC_OK;
}

static Csr csr_of(const Shard& s) { return Csr{s.indptr, s.col, s.q, s.n_rows}; }

static int launch_fused(tsc_handle* h, Shard& s, bool gated);
static int launch_reassign_sums(tsc_handle* h, Shard& s, int method, double thresh, const double* ta, const double* tu,
                                double* colsum, int* nbest_d);
static int launch_lnl_kernels(tsc_handle* h, Shard& s, const EmState* st, const double* ta, const double* tu,
                              const double* ia, const double* iu, int* nparts_out);

// ------------------------------------------------------------------------------------------------- tile launches
// CTAs of a tile pass: the resident grid, or fewer when the tile list is short (one tile per warp at least)
static int tiles_grid(const Shard& s, long long n_tiles) {
    return (int)std::max<long long>(1, std::min<long long>(s.grid_tiles, (n_tiles + kTileWarps - 1) / kTileWarps));
}

static int build_tiles(tsc_handle* h, Shard& s, const long long* indptr_d, long long n_rows, Tile** tiles_out,
                       long long* n_tiles_out, long long* n_long_out, struct Arena* arena, cudaStream_t st);
// the shard-wide tile list, built on first use (see create_attempt)
static int ensure_tiles(tsc_handle* h, Shard& s) {
    if (s.tiles || s.n_rows == 0) return TSC_OK;
    CU(cudaSetDevice(s.dev));
    return build_tiles(h, s, s.indptr, s.n_rows, &s.tiles, &s.n_tiles, &s.n_long, nullptr, nullptr);
}

template <int MODE>
static void launch_tiles(const Shard& s, const TileArgs& a, bool smem_tab, int long8_override = -1) {
    const size_t scratch = sizeof(double) * kTileWarps * kScratch;
    // > 0.5 % of the tiles are long reads (the residual CSR of the ELL path passes its own ratio)
    const bool long8 = long8_override >= 0 ? long8_override != 0 : s.n_long * 200 > s.n_tiles;
    const int grid = tiles_grid(s, a.n_tiles);
    if (MODE == TILE_FUSED && smem_tab) {
        if (long8) k_tiles<MODE, true, true><<<grid, kTileThreads, s.smem_tiles, s.stream>>>(a);
        else k_tiles<MODE, true, false><<<grid, kTileThreads, s.smem_tiles, s.stream>>>(a);
    } else {
        if (long8) k_tiles<MODE, false, true><<<grid, kTileThreads, scratch, s.stream>>>(a);
        else k_tiles<MODE, false, false><<<grid, kTileThreads, scratch, s.stream>>>(a);
    }
}

// reassign(method, thresh) column sums (into colsum, already zeroed; may be NULL) and / or best-hit counts per read
// (nbest_d, may be NULL) of one shard, posterior from the tables (ta, tu).  Clustered-stream layout: the stream kernel
// for its reads (lane = read: max / count / kept are private reductions) + the rows kernel on the residual CSR;
// otherwise the rows kernel on the whole shard.  Not for TSC_CHOOSE with picks or per-entry output.
static int launch_reassign_sums(tsc_handle* h, Shard& s, int method, double thresh, const double* ta, const double* tu,
                                double* colsum, int* nbest_d) {
    if (h->kernel == TSC_KERNEL_ELL) {
        const bool wanted = (colsum || nbest_d) && (method != TSC_UNIQUE || nbest_d);
        if (s.ell_slices > 0 && wanted) {
            EllArgs e{s.ell_stream, s.ell_index, s.ell_range, s.ell_slices, ta, colsum, h->K, 1, nullptr, nullptr, nullptr, nullptr,
                      method, thresh, s.ell_rowid, nbest_d};
            k_ell<ELL_REASSIGN><<<s.ell_grid, 32, ell_smem_bytes<ELL_REASSIGN>(), s.stream>>>(e);
            LAUNCH(h);
        }
        if (s.ell_long > 0 && wanted) {
            EllArgs e{s.ell_stream, s.ell_index + s.ell_slices, s.ell_lrange, s.ell_long, ta, colsum, h->K, 1, nullptr, nullptr, nullptr, nullptr,
                      method, thresh, s.ell_rowid + s.ell_slices * kEllReads, nbest_d};
            k_ell_long<ELL_REASSIGN><<<s.ell_lgrid, 32, ell_long_smem_bytes<ELL_REASSIGN>(), s.stream>>>(e);
            LAUNCH(h);
        }
        if (s.res_rows > 0) {
            ReassignArgs g{method, thresh, nullptr, nbest_d, colsum, nullptr, s.res_rowid};
            const Csr rest{s.res_indptr, s.res_col, s.res_q, s.res_rows};
            launch_rows(h->G, [&](auto gg) {
                k_reassign_rows<decltype(gg)::value><<<s.grid_rows, 512, 0, s.stream>>>(rest, ta, tu, g);
            });
            LAUNCH(h);
        }
    } else {
        ReassignArgs g{method, thresh, nullptr, nbest_d, colsum, nullptr, nullptr};
        launch_rows(h->G, [&](auto gg) {
            k_reassign_rows<decltype(gg)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), ta, tu, g);
        });
        LAUNCH(h);
    }
    CU(cudaGetLastError());
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- small API
extern "C" int tsc_abi_version(void) { return TSC_ABI_VERSION; }
extern "C" const char* tsc_last_error(void) { return g_err.c_str(); }

extern "C" int tsc_device_count(int32_t* n_out) {
    if (!n_out) return fail(TSC_ERR_ARG, "n_out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *n_out = 0; return fail(TSC_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e)); }
    *n_out = n;
    return TSC_OK;
}

extern "C" int tsc_set_nccl_path(const char* path) {
    g_nccl_path = path ? path : "";
    return TSC_OK;
}

extern "C" int tsc_nccl_unique_id(void* out128) {
    if (!out128) return fail(TSC_ERR_ARG, "out128 is NULL");
    int rc = nccl_load();
    if (rc) return rc;
    ncclUniqueId id;
    NC(g_nccl.GetUniqueId(&id));
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    memcpy(out128, &id, 128);
    return TSC_OK;
}

extern "C" void tsc_config_default(tsc_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof(*cfg));
    cfg->n_local_devices = 1;
    cfg->n_procs = 1;
    cfg->kernel = TSC_KERNEL_AUTO;
    cfg->replicas = 0;
    cfg->smem_table_cols = -1;
    cfg->smem_acc_cols = -1;
    cfg->permute_columns = 0;
    cfg->transport = TSC_TRANSPORT_AUTO;
}

static void peer_buffer_release(void* buf);

static void free_shard(Shard& s) {
    cudaSetDevice(s.dev);
    if (s.stream) cudaStreamSynchronize(s.stream);
    if (s.aux) cudaStreamSynchronize(s.aux);
    void* ptrs[] = {s.indptr, s.col, s.q, s.wy, s.tiles, s.pi, s.theta, s.pt, s.pi_prev, s.theta_prev, s.pt_prev,
                    s.pi_init, s.theta_init, s.pisum0, s.acc, s.thetasum, s.ones, s.tmp_a, s.tmp_b, s.tmp_c, s.colsum,
                    s.perm, s.rep, s.consts, s.st, s.diffs, s.lnls, s.partials, s.scalars, s.bad,
                    s.ell_stream, s.ell_index, s.ell_range, s.ell_range_lnl, s.ell_cta_ns, s.ell_rebal, s.ell_lrange, s.ell_lrange_lnl, s.ell_rowid, s.res_indptr, s.res_col, s.res_q, s.res_wy, s.res_tiles, s.res_tiles_amb,
                    s.peer_ptrs_d, s.tail_partials, s.tail_ticket, s.peer_err, s.log_tab};
    for (void* p : ptrs) if (p && !s.in_slab(p)) dev_free(p);
    for (auto& b : s.slabs) dev_free(b.first);
    // (mappings of other processes' buffers stay open in g_ipc for the next model)
    peer_buffer_release(s.peer_own);
    pinned_word_give(s.st_host);
    for (auto& e : s.ev_poll) if (e) cudaEventDestroy(e);
    for (auto& e : s.ev_k) if (e) cudaEventDestroy(e);
    for (auto& e : s.ev_em) if (e) cudaEventDestroy(e);
    if (s.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s.comm);
    if (s.ev_up) cudaEventDestroy(s.ev_up);
    if (s.aux) cudaStreamDestroy(s.aux);
    if (s.stream) cudaStreamDestroy(s.stream);
    s = Shard();
}

extern "C" void tsc_destroy(tsc_handle* h) {
    if (!h) return;
    if (h->peer_thread.joinable()) h->peer_thread.join();
    for (auto& s : h->shards) free_shard(s);
    delete h;
}


// ------------------------------------------------------------------------------------------------- transport
static void peer_geometry(int K, int* kpad, int* nb, int* cap) {
    *kpad = (K + 31) & ~31;
    *nb = (K + kTailBlockLoci - 1) / kTailBlockLoci;
    *cap = *kpad + 64;          // larger reductions go through the inbox in pieces
}

// Exchange buffers that other processes map (one process per GPU) and the mappings of theirs are kept for the life of
// the process: cudaIpcOpenMemHandle costs milliseconds per peer -- tens when eight processes do it at once -- and a
// process that builds model after model (assign + resume, a service, the benchmark's repeated jobs) meets the same peers
// with the same buffers every time.  A destroyed model hands its exported buffer back to `idle`; the next model of the same
// geometry re-exports it under the same handle, and its peers find the handle among the mappings they kept open.
struct IpcCache {
    std::mutex mu;
    struct Own { int dev; size_t bytes; unsigned char* buf; cudaIpcMemHandle_t hd; };
    std::vector<Own> idle, live;
    std::map<std::string, void*> opened;          // 64 handle bytes -> mapping in this process
};
static IpcCache g_ipc;

static int peer_buffer_alloc(int dev, int K, int world, unsigned char** out, bool ipc) {
    int kpad, nb, cap;
    peer_geometry(K, &kpad, &nb, &cap);
    const size_t bytes = 8 * peer_buffer_words(world, kpad, nb, cap);
    CU(cudaSetDevice(dev));
    if (ipc) CU(cudaMalloc(out, bytes));         // other processes map it: a block of its own, returned to the driver
    else CU(dev_malloc(out, bytes));
    CU(cudaMemset(*out, 0, bytes));            // flags start at epoch 0, before anybody can map the buffer
    CU(cudaDeviceSynchronize());
    return TSC_OK;
}

extern "C" int tsc_peer_buffer_create(int32_t device, int32_t n_cols, int32_t world, void** buf_out, void* ipc_handle64_out) {
    if (!buf_out || !ipc_handle64_out || n_cols <= 0 || world <= 0) return fail(TSC_ERR_ARG, "bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    int kpad, nb, cap;
    peer_geometry(n_cols, &kpad, &nb, &cap);
    const size_t bytes = 8 * peer_buffer_words(world, kpad, nb, cap);
    {   // a buffer of this geometry that an earlier model of this process exported: same handle, flags back to epoch 0
        std::unique_lock<std::mutex> g(g_ipc.mu);
        for (size_t i = 0; i < g_ipc.idle.size(); ++i) {
            if (g_ipc.idle[i].dev != device || g_ipc.idle[i].bytes != bytes) continue;
            const IpcCache::Own o = g_ipc.idle[i];
            g_ipc.idle.erase(g_ipc.idle.begin() + i);
            g_ipc.live.push_back(o);
            g.unlock();
            CU(cudaSetDevice(device));
            CU(cudaMemset(o.buf, 0, bytes));
            CU(cudaDeviceSynchronize());
            memcpy(ipc_handle64_out, &o.hd, 64);
            *buf_out = o.buf;
            return TSC_OK;
        }
    }
    unsigned char* buf = nullptr;
    int rc = peer_buffer_alloc(device, n_cols, world, &buf, true);
    if (rc) return rc;
    cudaIpcMemHandle_t hd;
    cudaError_t e = cudaIpcGetMemHandle(&hd, buf);
    if (e != cudaSuccess) { cudaFree(buf); return fail(TSC_ERR_CUDA, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); }
    memcpy(ipc_handle64_out, &hd, 64);
    *buf_out = buf;
    std::lock_guard<std::mutex> g(g_ipc.mu);
    g_ipc.live.push_back(IpcCache::Own{device, bytes, buf, hd});
    return TSC_OK;
}

// An exported buffer goes back to the process's pool (its peers may keep it mapped); anything else is freed.
static void peer_buffer_release(void* buf) {
    if (!buf) return;
    {
        std::lock_guard<std::mutex> g(g_ipc.mu);
        for (size_t i = 0; i < g_ipc.live.size(); ++i) {
            if (g_ipc.live[i].buf != buf) continue;
            g_ipc.idle.push_back(g_ipc.live[i]);
            g_ipc.live.erase(g_ipc.live.begin() + i);
            return;
        }
    }
    dev_free(buf);
}

extern "C" void tsc_peer_buffer_free(void* buf) { peer_buffer_release(buf); }

// Decide how the shards exchange K-vectors and set it up:
//   peer  - every GPU maps every rank's exchange buffer (same process: peer access; one process per GPU: CUDA IPC
//           handles passed in cfg.peer_handles).  Also used with a single GPU (its own buffer), so that the
//           per-iteration tail is one kernel everywhere.
//   nccl  - ncclCommInitRank + ncclAllReduce (several nodes' worth of generality; slower to set up).
static int setup_transport(tsc_handle* h, const tsc_config& cfg) {
    const int n_local = (int)h->shards.size();
    peer_geometry(h->K, &h->kpad, &h->nb_tail, &h->cap);
    int want = cfg.transport;
    if (want != TSC_TRANSPORT_PEER && want != TSC_TRANSPORT_NCCL) {            // auto
        want = TSC_TRANSPORT_PEER;
        if (h->n_procs > 1 && !cfg.peer_handles) want = TSC_TRANSPORT_NCCL;
        if (h->n_procs > 1 && n_local > 1) want = TSC_TRANSPORT_NCCL;
        if (h->n_procs == 1 && n_local > 1) {
            for (int i = 0; i < n_local && want == TSC_TRANSPORT_PEER; ++i)
                for (int j = 0; j < n_local; ++j) {
                    int can = 1;
                    if (i != j) CU(cudaDeviceCanAccessPeer(&can, h->shards[i].dev, h->shards[j].dev));
                    if (!can) { want = TSC_TRANSPORT_NCCL; break; }
                }
        }
    }
    h->transport = want;
    if (want == TSC_TRANSPORT_NCCL) {
        if (h->world == 1) { h->transport = TSC_TRANSPORT_PEER; want = TSC_TRANSPORT_PEER; }
    }
    if (want == TSC_TRANSPORT_NCCL) {
        ncclUniqueId id;
        int rc = nccl_load();
        if (rc) return rc;
        if (h->n_procs > 1) {
            if (!cfg.nccl_id) return fail(TSC_ERR_ARG, "the NCCL transport over several processes needs nccl_id");
            memcpy(&id, cfg.nccl_id, 128);
        } else NC(g_nccl.GetUniqueId(&id));
        NC(g_nccl.GroupStart());
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            NC(g_nccl.CommInitRank(&s.comm, h->world, id, s.world_rank));
        }
        NC(g_nccl.GroupEnd());
    } else {
        if (h->n_procs > 1 && n_local > 1) return fail(TSC_ERR_ARG, "the peer transport takes one GPU per process or one process");
        if (h->n_procs > 1 && (!cfg.peer_buffer || !cfg.peer_handles))
            return fail(TSC_ERR_ARG, "the peer transport over several processes needs peer_buffer and peer_handles");
        for (auto& s : h->shards) {
            s.peer_map.assign(h->world, nullptr);
            s.peer_opened.assign(h->world, 0);
            if (h->n_procs > 1) s.peer_own = (unsigned char*)cfg.peer_buffer;           // ownership moves to the handle
            else { int rc = peer_buffer_alloc(s.dev, h->K, h->world, &s.peer_own, false); if (rc) return rc; }
            s.peer_map[s.world_rank] = s.peer_own;
        }
        if (h->n_procs > 1) {
            // cudaIpcOpenMemHandle costs ~10 ms per peer: it runs beside the upload of the CSR; peer_ready() joins
            {
                Shard& s0 = h->shards[0];
                CU(cudaSetDevice(s0.dev));
                CU(dev_malloc(&s0.peer_ptrs_d, sizeof(unsigned char*) * h->world));
            }
            std::vector<char> handles((const char*)cfg.peer_handles, (const char*)cfg.peer_handles + 64 * (size_t)h->world);
            h->peer_thread = std::thread([h, handles]() {
                Shard& s = h->shards[0];
                auto bad = [&](const std::string& what, cudaError_t e) {
                    h->peer_thread_rc = TSC_ERR_CUDA;
                    h->peer_thread_err = what + ": " + cudaGetErrorString(e);
                };
                cudaError_t e = cudaSetDevice(s.dev);
                if (e != cudaSuccess) return bad("cudaSetDevice", e);
                for (int r = 0; r < h->world; ++r) {
                    if (r == s.world_rank) continue;
                    cudaIpcMemHandle_t hd;
                    memcpy(&hd, handles.data() + 64 * (size_t)r, 64);
                    void* p = nullptr;
                    const std::string hkey((const char*)&hd, 64);
                    {
                        std::lock_guard<std::mutex> g(g_ipc.mu);
                        auto it = g_ipc.opened.find(hkey);
                        if (it != g_ipc.opened.end()) p = it->second;
                    }
                    if (p) { s.peer_map[r] = (unsigned char*)p; continue; }
                    e = cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess);
                    if (e != cudaSuccess)
                        return bad("cudaIpcOpenMemHandle(rank " + std::to_string(r) + ") -- the peer transport needs every GPU of "
                                   "the node visible to every rank", e);
                    s.peer_map[r] = (unsigned char*)p;
                    s.peer_opened[r] = 1;
                    { std::lock_guard<std::mutex> g(g_ipc.mu); g_ipc.opened[hkey] = p; }
                }
                e = cudaMemcpy(s.peer_ptrs_d, s.peer_map.data(), sizeof(unsigned char*) * h->world, cudaMemcpyHostToDevice);
                if (e != cudaSuccess) return bad("peer pointer table", e);
            });
        } else if (n_local > 1) {
            for (auto& a : h->shards) {
                CU(cudaSetDevice(a.dev));
                for (auto& b : h->shards) {
                    if (&a == &b) continue;
                    cudaError_t e = cudaDeviceEnablePeerAccess(b.dev, 0);
                    if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
                    if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e));
                    a.peer_map[b.world_rank] = b.peer_own;
                }
            }
        }
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            if (h->n_procs == 1) {
                CU(dev_malloc(&s.peer_ptrs_d, sizeof(unsigned char*) * h->world));
                CU(cudaMemcpy(s.peer_ptrs_d, s.peer_map.data(), sizeof(unsigned char*) * h->world, cudaMemcpyHostToDevice));
            }
            CU(dev_malloc(&s.tail_partials, sizeof(double) * h->nb_tail));
            CU(dev_malloc(&s.tail_ticket, sizeof(unsigned)));
            CU(cudaMemset(s.tail_ticket, 0, sizeof(unsigned)));
            CU(dev_malloc(&s.peer_err, sizeof(int)));
            CU(cudaMemset(s.peer_err, 0, sizeof(int)));
        }
    }
    return TSC_OK;
}


// ------------------------------------------------------------------------------------------------- host -> device
// cudaMemcpyAsync from ordinary (pageable) memory goes through the driver's single staging buffer at ~11 GB/s.  A
// scipy caller hands over exactly such arrays, so large pageable sources are staged here instead: a few host threads
// copy chunks into a process-wide pool of page-locked slots and queue the DMA of each chunk as soon as it is filled
// (~40 GB/s on the 16-core host of a B200 box, against ~55 GB/s from memory that is already page-locked).
struct StagePool {
    static constexpr size_t kSlot = 8u << 20;
    std::vector<void*> slots;
    std::mutex mu;
    int ensure(size_t n) {
        std::lock_guard<std::mutex> g(mu);
        while (slots.size() < n) {
            void* p = nullptr;
            if (cudaHostAlloc(&p, kSlot, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); break; }
            slots.push_back(p);
        }
        return (int)slots.size();
    }
};
static StagePool g_stage;
static std::mutex g_stage_busy;        // one staged upload at a time per process (the slots are shared)

static bool host_is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

// dst (device, on s.dev) <- src (host), in stream order on s.stream.  Returns once every byte has left `src`'s pages or
// (pinned source) once the copy is queued.
static int upload(tsc_handle* h, Shard& s, void* dst, const void* src, size_t bytes) {
    if (bytes == 0) return TSC_OK;
    h->h2d += (long long)bytes;
    int threads = (int)std::thread::hardware_concurrency() / std::max(1, h->n_procs * (int)h->shards.size());
    threads = std::max(1, std::min(threads, 8));               // (12 threads measured slower than 8 on the 16-core host)
    if (bytes < (32u << 20) || host_is_pinned(src) || getenv("TELESCOPE_B200_NO_STAGING")) {
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.stream));
        return TSC_OK;
    }
    std::lock_guard<std::mutex> busy(g_stage_busy);
    const int n_slots = g_stage.ensure((size_t)2 * threads);
    if (n_slots < 2) {      // no page-locked memory to be had: the plain path still works
        CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s.stream));
        return TSC_OK;
    }
    threads = std::min(threads, n_slots / 2);
    const int slots_used = 2 * threads;                 // a slot is only ever touched by one thread
    const size_t n_chunks = (bytes + StagePool::kSlot - 1) / StagePool::kSlot;
    std::vector<cudaError_t> err(threads, cudaSuccess);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t]() {
            cudaError_t e = cudaSetDevice(s.dev);
            cudaEvent_t ev[2] = {nullptr, nullptr};
            for (int k = 0; k < 2 && e == cudaSuccess; ++k) e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
            bool used[2] = {false, false};
            for (size_t i = t; i < n_chunks && e == cudaSuccess; i += threads) {
                const int k = (int)((i / threads) & 1);                  // this thread's two slots alternate
                void* slot = g_stage.slots[(size_t)t + (size_t)k * threads];
                if (used[k]) e = cudaEventSynchronize(ev[k]);            // the DMA out of this slot has finished
                if (e != cudaSuccess) break;
                const size_t off = i * StagePool::kSlot, n = std::min(StagePool::kSlot, bytes - off);
                memcpy(slot, (const char*)src + off, n);
                e = cudaMemcpyAsync((char*)dst + off, slot, n, cudaMemcpyHostToDevice, s.stream);
                if (e == cudaSuccess) e = cudaEventRecord(ev[k], s.stream);
                used[k] = true;
            }
            for (int k = 0; k < 2; ++k) {
                if (used[k] && e == cudaSuccess) e = cudaEventSynchronize(ev[k]);     // slots go back to the pool drained
                if (ev[k]) cudaEventDestroy(ev[k]);
            }
            err[t] = e;
        });
    }
    for (auto& th : pool) th.join();
    (void)slots_used;
    for (cudaError_t e : err)
        if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("staged upload: ") + cudaGetErrorString(e));
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- create
struct StageTimer {
    bool on;                 // TELESCOPE_B200_TIMING: print laps and synchronise at lap boundaries
    std::string* log;        // always: host-side laps, no extra synchronisation
    std::chrono::steady_clock::time_point t0;
    explicit StageTimer(std::string* l) : on(getenv("TELESCOPE_B200_TIMING") != nullptr), log(l), t0(std::chrono::steady_clock::now()) {}
    void lap(const char* what) {
        auto t1 = std::chrono::steady_clock::now();
        const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        if (on) fprintf(stderr, "[tsc_create] %-28s %8.1f ms\n", what, ms);
        if (log) { char buf[96]; snprintf(buf, sizeof buf, "%s=%.1f;", what, ms); *log += buf; }
        t0 = t1;
    }
};

struct CreateInput {
    int64_t n_rows_user;
    int32_t n_cols;
    int64_t nnz;
    const void* indptr;
    int32_t indptr_bytes;
    const int32_t* indices;
    const uint16_t* raw;
    const double* q_lut;
    int32_t lut_len;
    double pi_prior, theta_prior;
    long long ip_at(int64_t i) const {
        return indptr_bytes == 4 ? (long long)((const int32_t*)indptr)[i] : (long long)((const int64_t*)indptr)[i];
    }
};

constexpr int kNeedsCompaction = -1;    // internal: empty reads were found on the device, retry with compacted reads

// Reads with no entries are dropped on the host (slow path): ip becomes the compacted read pointers, h->rowmap the
// compacted -> caller read index.
static int compact_on_host(tsc_handle* h, const CreateInput& in, std::vector<long long>& ip) {
    std::vector<long long> full((size_t)in.n_rows_user + 1);
    for (int64_t i = 0; i <= in.n_rows_user; ++i) full[i] = in.ip_at(i);
    for (int64_t i = 0; i < in.n_rows_user; ++i)
        if (full[i + 1] < full[i]) return fail(TSC_ERR_ARG, "indptr is not non-decreasing");
    h->rowmap.clear();
    h->rowmap.reserve(in.n_rows_user);
    ip.clear();
    ip.reserve(in.n_rows_user + 1);
    ip.push_back(0);
    for (int64_t i = 0; i < in.n_rows_user; ++i)
        if (full[i + 1] > full[i]) { h->rowmap.push_back(i); ip.push_back(full[i + 1]); }
    return TSC_OK;
}

static int create_attempt(tsc_handle* h, const tsc_config& cfg, const CreateInput& in, const std::vector<long long>* ip_compact,
                          StageTimer& tm);

// Flat tiles of whole reads (tsc_tiles.cuh) over a device-resident read-pointer array: count per chunk, prefix on the
// host (a few thousand chunks), fill.
struct Arena {               // bump allocator over device memory that is already there (dead upload buffers)
    char* base = nullptr;
    size_t cap = 0, off = 0;
    void* take(size_t bytes) {
        const size_t a = (off + 255) & ~(size_t)255;
        if (!base || a + bytes > cap) return nullptr;
        off = a + bytes;
        return base + a;
    }
};

struct DevBuf {              // device temporaries released on every exit path
    std::vector<void*> p;
    Arena* arena = nullptr;  // tried first; cudaMalloc only when it is full (every cudaMalloc/cudaFree of a large
                             // block costs milliseconds and a device synchronisation)
    ~DevBuf() { for (void* q : p) if (q) dev_free(q); }
    template <typename T> cudaError_t alloc(T** out, size_t n) {
        const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
        if (arena) { void* q = arena->take(bytes); if (q) { *out = (T*)q; return cudaSuccess; } }
        void* q = nullptr;
        cudaError_t e = dev_malloc(&q, bytes);
        *out = (T*)q;
        if (e == cudaSuccess) p.push_back(q);
        return e;
    }
    void keep(void* q) {     // q outlives the buffer: true when it was cudaMalloc'ed here (the caller now owns it)
        auto it = std::find(p.begin(), p.end(), q);
        if (it != p.end()) p.erase(it);
    }
    bool owns(void* q) const { return std::find(p.begin(), p.end(), q) != p.end(); }
};

struct SlabPlan {            // sizes first, one cudaMalloc, then the pointers
    struct Item { void** out; size_t bytes; };
    std::vector<Item> items;
    template <typename T> void add(T** out, size_t n) { items.push_back({(void**)out, std::max<size_t>(n, 1) * sizeof(T)}); }
    int commit(Shard& s) {
        size_t total = 0;
        for (auto& it : items) total += (it.bytes + 255) & ~(size_t)255;
        char* base = nullptr;
        CU(dev_malloc(&base, std::max<size_t>(total, 256)));
        s.slabs.push_back({base, std::max<size_t>(total, 256)});
        size_t off = 0;
        for (auto& it : items) { *it.out = base + off; off += (it.bytes + 255) & ~(size_t)255; }
        return TSC_OK;
    }
};

static int build_tiles(tsc_handle* h, Shard& s, const long long* indptr_d, long long n_rows, Tile** tiles_out,
                       long long* n_tiles_out, long long* n_long_out, Arena* arena, cudaStream_t st) {
    if (!st) st = s.stream;
    const int n_chunks = (int)((n_rows + kChunkRows - 1) / kChunkRows);
    DevBuf tmp;
    tmp.arena = arena;
    Arena mark;
    if (arena) mark = *arena;
    int* counts_d = nullptr;
    long long* offs_d = nullptr;
    unsigned long long* nlong_d = nullptr;
    CU(tmp.alloc(&counts_d, n_chunks));
    CU(tmp.alloc(&offs_d, n_chunks));
    CU(tmp.alloc(&nlong_d, 1));
    CU(cudaMemsetAsync(nlong_d, 0, sizeof(unsigned long long), st));
    std::vector<int> counts(n_chunks);
    std::vector<long long> offs(n_chunks);
    if (n_chunks > 0) {
        k_tile_count<<<(n_chunks + 127) / 128, 128, 0, st>>>(indptr_d, n_rows, counts_d, n_chunks);
        LAUNCH(h);
        CU(cudaGetLastError());
        CU(cudaMemcpyAsync(counts.data(), counts_d, sizeof(int) * n_chunks, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
    }
    long long total = 0;
    for (int c = 0; c < n_chunks; ++c) { offs[c] = total; total += counts[c]; }
    *n_tiles_out = total;
    CU(dev_malloc(tiles_out, sizeof(Tile) * std::max<long long>(total, 1)));
    if (n_chunks > 0) {
        CU(cudaMemcpyAsync(offs_d, offs.data(), sizeof(long long) * n_chunks, cudaMemcpyHostToDevice, st));
        k_tile_fill<<<(n_chunks + 127) / 128, 128, 0, st>>>(indptr_d, n_rows, offs_d, n_chunks, *tiles_out, nlong_d);
        LAUNCH(h);
        CU(cudaGetLastError());
    }
    unsigned long long nl = 0;
    CU(cudaMemcpyAsync(&nl, nlong_d, sizeof(nl), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    *n_long_out = (long long)nl;
    if (arena) arena->off = mark.off;
    return TSC_OK;
}

// out[0..n] = exclusive prefix sums of in[0..n) on the shard's stream (tsc_ell.cuh scan kernels)
template <typename T>
static int device_scan(tsc_handle* h, Shard& s, const T* in, long long n, long long* out, Arena* arena = nullptr) {
    const int nb = (int)std::max<long long>(1, (n + kScanItems - 1) / kScanItems);
    DevBuf tmp;
    Arena mark;
    if (arena) { mark = *arena; tmp.arena = arena; }
    long long* tot = nullptr;
    CU(tmp.alloc(&tot, (size_t)nb + 1));
    k_scan_local<T><<<nb, 1024, 0, s.stream>>>(in, n, out, tot);
    k_scan_totals<<<1, 1024, 0, s.stream>>>(tot, nb);
    k_scan_add<<<grid_for(n, 256, s.n_sm * 16), 256, 0, s.stream>>>(out, n, tot, nb);
    h->launches += 3;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(s.stream));      // tot is released on return
    if (arena) arena->off = mark.off;
    return TSC_OK;
}

static int fetch_ll(Shard& s, const long long* dev, long long* host) {
    CU(cudaMemcpyAsync(host, dev, sizeof(long long), cudaMemcpyDeviceToHost, s.stream));
    CU(cudaStreamSynchronize(s.stream));
    return TSC_OK;
}

constexpr bool kLongRecordsDefault = true;     // long-read records in the stream (k_ell_long); see DESIGN.md

// The clustered sliced-ELL stream of the fused kernel and the residual CSR (tsc_ell.cuh), from the shard's finished
// q / col / indptr / wy arrays.  One-off; what it keeps belongs to the shard.  `arena`: dead device memory (the raw
// upload buffers) that serves the temporaries.
static int build_ell(tsc_handle* h, Shard& s, StageTimer& tm, Arena* arena) {
    auto lap = [&](const char* what) { if (tm.on) cudaStreamSynchronize(s.stream); tm.lap(what); };
    const int K = h->K;
    const long long n_rows = s.n_rows;
    CU(cudaSetDevice(s.dev));
    DevBuf tmp;
    tmp.arena = arena;
    int* key = nullptr;
    // [0] ambiguous reads, [1] their entries, [2] reads outside the stream (packed count | entries), [3] the ambiguous
    // ones among them (packed), [4] append cursor
    unsigned long long* counters = nullptr;
    CU(tmp.alloc(&counters, 8));
    CU(cudaMemsetAsync(counters, 0, sizeof(unsigned long long) * 8, s.stream));
    const long long n_short_keys_ll = (long long)K << kEllLenBits, n_keys_ll = n_short_keys_ll + K;
    const bool ell_ok = n_rows > 0 && n_keys_ll < (1LL << 31);
    long long n_short = 0, n_long = 0;
    int n_stream_keys = 0x7fffffff;             // keys at or above this are not in the stream after all
    int* sorted = nullptr;
    long long slots_short = 0;
    if (ell_ok) {
        const int n_short_keys = (int)n_short_keys_ll, n_keys = (int)n_keys_ll;
        unsigned *hist = nullptr, *cursor = nullptr;
        long long* bin_start = nullptr;
        CU(tmp.alloc(&key, n_rows));
        CU(tmp.alloc(&hist, n_keys));
        CU(tmp.alloc(&cursor, n_keys));
        CU(tmp.alloc(&bin_start, (size_t)n_keys + 1));
        CU(cudaMemsetAsync(hist, 0, sizeof(unsigned) * n_keys, s.stream));
        CU(cudaMemsetAsync(cursor, 0, sizeof(unsigned) * n_keys, s.stream));
        const int g = grid_for(n_rows, 256, s.n_sm * 16);
        k_ell_classify<<<g, 256, 0, s.stream>>>(s.indptr, n_rows, s.col, n_short_keys, n_keys, key, hist);
        LAUNCH(h);
        CU(cudaGetLastError());
        int rc = device_scan<unsigned>(h, s, hist, n_keys, bin_start, arena);
        if (rc) return rc;
        long long n_cand = 0;
        if ((rc = fetch_ll(s, bin_start + n_short_keys, &n_short))) return rc;
        if ((rc = fetch_ll(s, bin_start + n_keys, &n_cand))) return rc;
        n_long = n_cand - n_short;
        // a handful of long reads is not worth a launch per pass: they stay with the flat tiles of the residual
        const char* lr_env = getenv("TELESCOPE_B200_LONG_RECORDS");      // 0 / 1 overrides the default
        const bool long_records = lr_env ? atoi(lr_env) != 0 : kLongRecordsDefault;
        if (!long_records || n_long * 1000 < n_rows) { n_long = 0; n_cand = n_short; n_stream_keys = n_short_keys; }
        lap("  ell classify+hist+scan");
        if (n_cand > 0) {
            // kept: slot -> read of every slice, then the long reads (best-hit counts of reassign go back through it);
            // the short slots are padded to whole slices
            slots_short = ((n_short + kEllReads - 1) / kEllReads) * kEllReads;
            CU(dev_malloc(&s.ell_rowid, sizeof(int) * (slots_short + n_long)));
            CU(cudaMemsetAsync(s.ell_rowid, 0xff, sizeof(int) * (slots_short + n_long), s.stream));
            sorted = s.ell_rowid;
            k_ell_scatter<<<g, 256, 0, s.stream>>>(key, n_rows, bin_start, cursor, sorted, n_short_keys, n_stream_keys, slots_short - n_short);
            LAUNCH(h);
            CU(cudaGetLastError());
        }
    }
    if (n_short + n_long > 0) {
        const long long n_slices = slots_short / kEllReads, n_records = n_slices + n_long;
        int* rec_bytes = nullptr;
        long long* rec_off = nullptr;
        CU(dev_malloc(&s.ell_index, sizeof(int4) * n_records));      // kept: the kernels' record index
        CU(tmp.alloc(&rec_bytes, n_records));
        CU(tmp.alloc(&rec_off, (size_t)n_records + 1));
        lap("  ell scatter");
        if (n_slices > 0) {
            k_ell_slices<<<grid_for(n_slices, 128, s.n_sm * 16), 128, 0, s.stream>>>(s.indptr, s.col, sorted, n_short, n_slices, key,
                                                                                    s.ell_index, rec_bytes);
            LAUNCH(h);
        }
        if (n_long > 0) {
            k_ell_long_index<<<grid_for(n_long, 128, s.n_sm * 16), 128, 0, s.stream>>>(s.indptr, s.col, sorted + slots_short, n_long,
                                                                                      s.ell_index + n_slices, rec_bytes + n_slices);
            LAUNCH(h);
        }
        CU(cudaGetLastError());
        int rc = device_scan<int>(h, s, rec_bytes, n_records, rec_off, arena);
        if (rc) return rc;
        long long total = 0;
        if ((rc = fetch_ll(s, rec_off + n_records, &total))) return rc;
        int per_sm = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ell<ELL_FUSED>, 32, kEllSmem));
        int per_sm_lnl = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_lnl, k_ell<ELL_LNL>, 32, ell_smem_bytes<ELL_LNL>()));
        const int warps = s.n_sm * std::max(per_sm, 1);
        lap("  ell slices+scan");
        if (total >= (1LL << 36)) return fail(TSC_ERR_ARG, "slice stream of one GPU exceeds 64 GB");
        CU(dev_malloc(&s.ell_stream, (size_t)total));
        lap("  ell stream malloc");
        if (n_slices > 0) {
            k_ell_fill<<<grid_for(n_slices * 32, kFillWarps * 32, s.n_sm * 32), kFillWarps * 32, 0, s.stream>>>(s.indptr, s.col, s.q, s.wy, sorted, n_short, n_slices,
                                                                                       s.ell_index, rec_off, s.ell_stream);
            LAUNCH(h);
        }
        if (n_long > 0) {
            k_ell_fill_long<<<grid_for(n_long * 32, 256, s.n_sm * 16), 256, 0, s.stream>>>(s.indptr, s.col, s.q, s.wy, sorted + slots_short,
                                                                                          n_long, s.ell_index + n_slices, rec_off + n_slices,
                                                                                          s.ell_stream);
            LAUNCH(h);
        }
        CU(cudaGetLastError());
        s.ell_bytes = total;
        s.ell_slice_bytes = total;
        if (n_long > 0 && n_slices > 0 && (rc = fetch_ll(s, rec_off + n_slices, &s.ell_slice_bytes))) return rc;
        s.ell_slices = n_slices;
        s.ell_long = n_long;
        s.ell_records = n_records;
        // every resident warp gets a contiguous, byte-balanced run (at least ~8 records each when the stream is short);
        // slices and long reads have a kernel (and a split) each
        auto split = [&](const long long* off, long long n, int cap, int* grid, long long** range) -> int {
            *grid = (int)std::max<long long>(1, std::min<long long>((n + 7) / 8, cap));
            CU(dev_malloc(range, sizeof(long long) * (*grid + 1)));
            k_ell_ranges<<<(*grid + 256) / 256, 256, 0, s.stream>>>(off, n, *grid, *range);
            LAUNCH(h);
            return TSC_OK;
        };
        if (n_slices > 0) {
            if ((rc = split(rec_off, n_slices, warps, &s.ell_grid, &s.ell_range))) return rc;
            // the first iterations of the model measure how long every run takes and move the boundaries accordingly
            const char* rb_env = getenv("TELESCOPE_B200_REBALANCE");
            s.ell_rebal_left = rb_env ? atoi(rb_env) : 4;
            if (s.ell_grid < 2 || s.ell_grid + 1 > 4096) s.ell_rebal_left = 0;
            if (s.ell_rebal_left > 0) {
                CU(dev_malloc(&s.ell_cta_ns, sizeof(unsigned) * s.ell_grid));
                CU(dev_malloc(&s.ell_rebal, sizeof(double) * 2 * (s.ell_grid + 1)));
                CU(cudaMemsetAsync(s.ell_cta_ns, 0, sizeof(unsigned) * s.ell_grid, s.stream));
            }
            if ((rc = split(rec_off, n_slices, s.n_sm * std::max(per_sm_lnl, 1), &s.ell_grid_lnl, &s.ell_range_lnl))) return rc;
        }
        if (n_long > 0) {
            int pl = 0, pl_lnl = 0;
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pl, k_ell_long<ELL_FUSED>, 32, ell_long_smem_bytes<ELL_FUSED>()));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&pl_lnl, k_ell_long<ELL_LNL>, 32, ell_long_smem_bytes<ELL_LNL>()));
            // (the long-read kernel deals batches of 32 records round-robin: no run table)
            const long long batches = (n_long + 31) / 32;
            s.ell_lgrid = (int)std::max<long long>(1, std::min<long long>(batches, (long long)s.n_sm * std::max(pl, 1)));
            s.ell_lgrid_lnl = (int)std::max<long long>(1, std::min<long long>(batches, (long long)s.n_sm * std::max(pl_lnl, 1)));
        }
        CU(cudaGetLastError());
        lap("  ell fill");
    }
    // ---- residual CSR: every read without a slot in the stream -- the unique reads (they add nothing to the M-step
    // sums but take part in the log-likelihood) and the ambiguous reads that do not fit a slice (key < 0; all of them
    // when the stream could not be built).  They are counted first and then appended in whatever order the atomics
    // give: the order of the residual's reads is irrelevant, only each read's entries stay together.
    {
        const int g = grid_for(n_rows, 256, s.n_sm * 16);
        k_res_count<<<g, 256, 0, s.stream>>>(s.indptr, n_rows, key, n_stream_keys, counters);
        LAUNCH(h);
        CU(cudaGetLastError());
        unsigned long long cnt[4];
        CU(cudaMemcpyAsync(cnt, counters, sizeof(cnt), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        const long long amb_rows = (long long)cnt[0], amb_nnz = (long long)cnt[1];
        s.res_rows = (long long)(cnt[2] >> kResShift);
        s.res_nnz = (long long)(cnt[2] & ((1ULL << kResShift) - 1ULL));
        s.res_amb_rows = (long long)(cnt[3] >> kResShift);
        s.res_amb_nnz = (long long)(cnt[3] & ((1ULL << kResShift) - 1ULL));
        s.ell_reads = amb_rows - s.res_amb_rows;
        s.ell_entries = amb_nnz - s.res_amb_nnz;
        lap("  residual count");
        if (s.res_rows >= (1LL << (64 - kResShift))) return fail(TSC_ERR_ARG, "too many reads outside the slice stream on one GPU");
        if (s.res_rows > 0) {
            const size_t pad = 256;
            {
                SlabPlan rs;
                rs.add(&s.res_indptr, (size_t)s.res_rows + 1);
                rs.add(&s.res_col, (size_t)s.res_nnz + pad);
                rs.add(&s.res_q, (size_t)s.res_nnz + pad);
                rs.add(&s.res_wy, (size_t)s.res_rows);
                rs.add(&s.res_rowid, (size_t)s.res_rows);
                int rc = rs.commit(s);
                if (rc) return rc;
            }
            CU(cudaMemsetAsync(s.res_col + s.res_nnz, 0, sizeof(int) * pad, s.stream));
            CU(cudaMemsetAsync(s.res_q + s.res_nnz, 0, sizeof(double) * pad, s.stream));
            {   // counters[4]: cursor of the ambiguous front, counters[5]: cursor of the unique reads behind it
                const unsigned long long start[2] = {0ULL, ((unsigned long long)s.res_amb_rows << kResShift) | (unsigned long long)s.res_amb_nnz};
                CU(cudaMemcpyAsync(counters + 4, start, sizeof(start), cudaMemcpyHostToDevice, s.stream));
            }
            k_res_append<<<grid_for(n_rows, 256, s.n_sm * 16), 256, 0, s.stream>>>(s.indptr, n_rows, s.col, s.q, s.wy, key, n_stream_keys, counters + 4,
                                                                                     s.res_indptr, s.res_col, s.res_q, s.res_wy, s.res_rowid);
            LAUNCH(h);
            CU(cudaMemcpyAsync(s.res_indptr + s.res_rows, &s.res_nnz, sizeof(long long), cudaMemcpyHostToDevice, s.stream));
            CU(cudaGetLastError());
            CU(cudaStreamSynchronize(s.stream));
            int rc = build_tiles(h, s, s.res_indptr, s.res_rows, &s.res_tiles, &s.res_n_tiles, &s.res_n_long, arena, nullptr);
            if (!rc && s.res_amb_rows > 0)
                rc = build_tiles(h, s, s.res_indptr, s.res_amb_rows, &s.res_tiles_amb, &s.res_amb_tiles, &s.res_amb_long, arena, nullptr);
            if (rc) return rc;
            lap("  residual copy+tiles");
        }
    }
    CU(cudaStreamSynchronize(s.stream));
    return TSC_OK;
}


static int create_impl(tsc_handle* h, const tsc_config& cfg, const CreateInput& in) {
    h->K = in.n_cols;
    h->n_rows_user = in.n_rows_user;
    h->nnz = in.nnz;
    h->n_procs = std::max(1, cfg.n_procs);
    h->proc_rank = cfg.proc_rank;
    const int n_local = std::max(1, cfg.n_local_devices);
    h->world = h->n_procs * n_local;
    if (h->n_procs > 1 && !cfg.nccl_id && !cfg.peer_handles)
        return fail(TSC_ERR_ARG, "n_procs > 1 needs peer_handles (peer transport) or nccl_id (NCCL transport)");
    if (h->proc_rank < 0 || h->proc_rank >= h->n_procs) return fail(TSC_ERR_ARG, "proc_rank out of range");

    int ndev = 0;
    {
        cudaError_t e = cudaGetDeviceCount(&ndev);
        if (e != cudaSuccess || ndev == 0)
            return fail(TSC_ERR_CUDA, std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                                          "); telescope_b200 has no CPU fallback");
    }
    for (int i = 0; i < n_local; ++i) {
        const int d = cfg.device_ids ? cfg.device_ids[i] : i;
        if (d < 0 || d >= ndev) return fail(TSC_ERR_ARG, "device id " + std::to_string(d) + " not present");
    }
    if (in.ip_at(0) != 0 || in.ip_at(in.n_rows_user) != in.nnz)
        return fail(TSC_ERR_ARG, "indptr[0] must be 0 and indptr[n_rows] must be nnz");

    StageTimer tm(&h->create_laps);
    struct AllocStats {          // what this construction spent in the driver's allocation calls, appended to the laps
        std::string* log; double ms0; long long calls0, hits0;
        ~AllocStats() {
            std::lock_guard<std::mutex> g(g_cache.mu);
            char buf[128];
            snprintf(buf, sizeof buf, "driver alloc calls=%lld (%.1f ms), cache hits=%lld;", g_cache.calls - calls0, g_cache.ms - ms0,
                     g_cache.hits - hits0);
            *log += buf;
        }
    } alloc_stats{&h->create_laps, 0.0, 0, 0};
    { std::lock_guard<std::mutex> g(g_cache.mu); alloc_stats.ms0 = g_cache.ms; alloc_stats.calls0 = g_cache.calls; alloc_stats.hits0 = g_cache.hits; }
    // Fast path: the caller's read pointers are used as they are (validated and rebased on the device).  Matrices
    // with empty reads take the slow path: tiny ones are checked here, large ones are detected on the device and the
    // attempt is repeated once with the reads compacted on the host (streams and communicators are kept).
    std::vector<long long> ip;
    bool slow = false;
    if (in.n_rows_user > 0 && in.n_rows_user <= 4096)
        for (int64_t i = 0; i < in.n_rows_user && !slow; ++i) slow = in.ip_at(i + 1) <= in.ip_at(i);
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (slow) { int rc = compact_on_host(h, in, ip); if (rc) return rc; }
        const int rc = create_attempt(h, cfg, in, slow ? &ip : nullptr, tm);
        if (rc != kNeedsCompaction) return rc;
        if (slow) return fail(TSC_ERR_STATE, "internal: empty read after compaction");
        slow = true;
    }
    return fail(TSC_ERR_STATE, "internal: construction did not settle");
}

static int create_attempt(tsc_handle* h, const tsc_config& cfg, const CreateInput& in, const std::vector<long long>* ip_compact,
                          StageTimer& tm) {
    const int K = in.n_cols;
    const int n_local = std::max(1, cfg.n_local_devices);
    const bool slow = ip_compact != nullptr;
    const int64_t nnz = in.nnz;
    const int32_t indptr_bytes = in.indptr_bytes;
    const void* indptr = in.indptr;
    const int32_t* indices = in.indices;
    const uint16_t* raw = in.raw;
    const double* q_lut = in.q_lut;
    const int32_t lut_len = in.lut_len;
    const double pi_prior = in.pi_prior, theta_prior = in.theta_prior;
    const long long n_rows = slow ? (long long)ip_compact->size() - 1 : (long long)in.n_rows_user;
    h->n_rows = n_rows;
    auto row_ptr = [&](long long r) -> long long { return slow ? (*ip_compact)[r] : in.ip_at(r); };
    tm.lap("indptr copy+validate");
    // ---- shard boundaries: contiguous, balanced by entry count
    std::vector<long long> rb(n_local + 1, 0);
    rb[n_local] = n_rows;
    for (int i = 1; i < n_local; ++i) {
        const long long target = nnz / n_local * i;
        long long lo = rb[i - 1], hi = n_rows;
        while (lo < hi) { const long long mid = (lo + hi) / 2; if (row_ptr(mid) < target) lo = mid + 1; else hi = mid; }
        rb[i] = lo;
    }

    if (h->shards.empty()) h->shards.resize(n_local);
    for (int i = 0; i < n_local; ++i) {
        Shard& s = h->shards[i];
        s.dev = cfg.device_ids ? cfg.device_ids[i] : i;
        s.world_rank = h->proc_rank * n_local + i;
        s.row_begin = rb[i];
        s.n_rows = rb[i + 1] - rb[i];
        s.nnz_begin = row_ptr(rb[i]);
        s.nnz = row_ptr(rb[i + 1]) - row_ptr(rb[i]);
        if (s.n_rows >= (1LL << 31)) return fail(TSC_ERR_ARG, "more than 2^31 reads on one GPU");
        CU(cudaSetDevice(s.dev));
        if (!s.stream) CU(cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
        if (!s.aux) CU(cudaStreamCreateWithFlags(&s.aux, cudaStreamNonBlocking));
        if (!s.ev_up) CU(cudaEventCreateWithFlags(&s.ev_up, cudaEventDisableTiming));
        CU(cudaDeviceGetAttribute(&s.n_sm, cudaDevAttrMultiProcessorCount, s.dev));
    }
    if (!h->shards[0].peer_own && !h->shards[0].comm) {
        int rc = setup_transport(h, cfg);
        if (rc) return rc;
    }

    tm.lap("streams + transport");
    // ---- tuning
    h->R = cfg.replicas > 0 ? std::min(cfg.replicas, 64) : 16;
    h->R_used = h->R;
    const double avg = n_rows ? (double)nnz / (double)n_rows : 1.0;
    h->G = avg <= 5.0 ? 4 : avg <= 24.0 ? 8 : avg <= 64.0 ? 16 : 32;
    h->kernel = (cfg.kernel == TSC_KERNEL_ROWS || cfg.kernel == TSC_KERNEL_TILES) ? cfg.kernel : TSC_KERNEL_ELL;

    // ---- upload + build per shard
    std::vector<uint16_t*> raw_d(n_local, nullptr);
    std::vector<int*> colin_d(n_local, nullptr);
    std::vector<unsigned long long*> cnt_d(n_local, nullptr);
    std::vector<double*> lut_d(n_local, nullptr);
    auto cleanup_tmp = [&]() {};            // every temporary of this function lives in a slab of its shard

    const size_t pad = 256;
    const bool permute = cfg.permute_columns != 0;
    for (int i = 0; i < n_local; ++i) {
        Shard& s = h->shards[i];
        CU(cudaSetDevice(s.dev));
        {
            // one slab for the entry arrays.  Without a locus renumbering the caller's loci are uploaded straight into
            // their final array; the raw scores stay (2 B per entry) and serve as scratch of the clustering afterwards
            SlabPlan big;
            big.add(&s.indptr, (size_t)s.n_rows + 1);
            big.add(&s.col, (size_t)s.nnz + pad);
            big.add(&s.q, (size_t)s.nnz + pad);
            big.add(&s.wy, (size_t)s.n_rows);
            big.add(&s.raw, (size_t)s.nnz);
            if (permute) big.add(&colin_d[i], (size_t)s.nnz);
            int rc = big.commit(s);
            if (rc) return rc;
            raw_d[i] = s.raw;
            if (!permute) colin_d[i] = s.col;
            // ... and one for everything K-sized
            SlabPlan small;
            const size_t kk = (size_t)K;
            double** kv[] = {&s.pi, &s.theta, &s.pt, &s.pi_prev, &s.theta_prev, &s.pt_prev, &s.pi_init, &s.theta_init,
                             &s.pisum0, &s.thetasum, &s.ones, &s.tmp_a, &s.tmp_b, &s.tmp_c, &s.colsum};
            for (double** p : kv) small.add(p, kk);
            small.add(&s.acc, kk * h->R);
            small.add(&s.perm, kk);
            small.add(&cnt_d[i], kk * 5);
            small.add(&lut_d[i], (size_t)lut_len);
            small.add(&s.bad, 1);
            small.add(&s.consts, 1);
            small.add(&s.st, 1);
            small.add(&s.scalars, 8);
            small.add(&s.partials, (size_t)s.n_sm * 96);
            small.add(&s.log_tab, (size_t)kLogTab);
            rc = small.commit(s);
            if (rc) return rc;
            CU(cudaMemsetAsync(s.slabs.back().first, 0, s.slabs.back().second, s.stream));
        }
        CU(cudaMemsetAsync(s.col + s.nnz, 0, sizeof(int) * pad, s.stream));
        CU(cudaMemsetAsync(s.q + s.nnz, 0, sizeof(double) * pad, s.stream));
        if (tm.on) cudaStreamSynchronize(s.stream);
        tm.lap("  cudaMalloc");
        // read pointers: native dtype up, int64 + rebased + validated on the device
        {
            const size_t ib = slow ? sizeof(long long) : (size_t)indptr_bytes;
            DevBuf ipbuf;
            char* ip_native = nullptr;
            // (the Q array is not written before the read pointers are through: it lends its memory, no allocation call)
            if (ib * (size_t)(s.n_rows + 1) <= sizeof(double) * (size_t)s.nnz) ip_native = (char*)s.q;
            else CU(ipbuf.alloc(&ip_native, ib * (s.n_rows + 1)));
            const char* src = slow ? (const char*)(ip_compact->data() + s.row_begin) : (const char*)indptr + ib * s.row_begin;
            { int rc = upload(h, s, ip_native, src, ib * (s.n_rows + 1)); if (rc) return rc; }
            const int g = grid_for(s.n_rows + 1, 256, s.n_sm * 16);
            if (ib == 4) k_indptr_prepare<int><<<g, 256, 0, s.stream>>>((const int*)ip_native, s.n_rows + 1, s.nnz_begin, s.indptr, s.bad);
            else k_indptr_prepare<long long><<<g, 256, 0, s.stream>>>((const long long*)ip_native, s.n_rows + 1, s.nnz_begin, s.indptr, s.bad);
            LAUNCH(h);
            CU(cudaGetLastError());
            int flags = 0;
            CU(cudaMemcpyAsync(&flags, s.bad, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            if (flags & 1) return fail(TSC_ERR_ARG, "indptr is not non-decreasing");
            if (flags & 2) {            // empty reads: release this attempt's arrays and ask for the slow path
                for (auto& sh : h->shards) {        // release this attempt's slabs; streams and the transport stay
                    cudaSetDevice(sh.dev);
                    if (sh.stream) cudaStreamSynchronize(sh.stream);
                    if (sh.tiles && !sh.in_slab(sh.tiles)) dev_free(sh.tiles);
                    for (auto& b : sh.slabs) dev_free(b.first);
                    sh.slabs.clear();
                    sh.indptr = nullptr; sh.col = nullptr; sh.q = nullptr; sh.wy = nullptr; sh.bad = nullptr; sh.tiles = nullptr;
                    sh.raw = nullptr;
                    double** kv[] = {&sh.pi, &sh.theta, &sh.pt, &sh.pi_prev, &sh.theta_prev, &sh.pt_prev, &sh.pi_init, &sh.theta_init,
                                     &sh.pisum0, &sh.thetasum, &sh.ones, &sh.tmp_a, &sh.tmp_b, &sh.tmp_c, &sh.colsum, &sh.acc,
                                     &sh.scalars, &sh.partials};
                    for (double** p : kv) *p = nullptr;
                    sh.perm = nullptr; sh.consts = nullptr; sh.st = nullptr; sh.log_tab = nullptr;
                }
                return kNeedsCompaction;
            }
            CU(cudaMemsetAsync(s.bad, 0, sizeof(int), s.stream));
        }
        tm.lap("  indptr up + prepare");
        { int rc = upload(h, s, lut_d[i], q_lut, sizeof(double) * lut_len); if (rc) return rc; }
        // tiles over the whole shard: only the flat-tile kernel family iterates with them; with the clustered stream they
        // serve nothing but the posterior export (estep / self.z) and are built when that is first asked for
        if (h->kernel == TSC_KERNEL_TILES) {
            int rc = build_tiles(h, s, s.indptr, s.n_rows, &s.tiles, &s.n_tiles, &s.n_long, nullptr, s.aux);
            if (rc) return rc;
        }
        tm.lap("  tiles");
        // The entry arrays go up in chunks of whole reads; as soon as a chunk has landed, the second stream runs its share
        // of the per-entry construction work -- the locus signatures, Q from the table, w*Y / totals / pisum0 -- while the
        // copy engine moves the next chunk.  (With a locus renumbering Q has to wait for the global counts: old order.)
        {
            const long long per_chunk = 64LL << 20;                                     // entries
            const int n_chunks = (int)std::max<long long>(1, std::min<long long>(32, (s.nnz + per_chunk - 1) / per_chunk));
            long long r0 = 0;
            for (int c = 0; c < n_chunks; ++c) {
                long long r1 = s.n_rows;
                if (c + 1 < n_chunks) {
                    const long long target = s.nnz_begin + s.nnz / n_chunks * (c + 1);
                    long long lo = r0, hi = s.n_rows;
                    while (lo < hi) { const long long mid = (lo + hi) / 2; if (row_ptr(s.row_begin + mid) < target) lo = mid + 1; else hi = mid; }
                    r1 = lo;
                }
                const long long e0 = row_ptr(s.row_begin + r0) - s.nnz_begin, e1 = row_ptr(s.row_begin + r1) - s.nnz_begin;
                if (e1 > e0) {
                    int rc = upload(h, s, raw_d[i] + e0, raw + s.nnz_begin + e0, sizeof(uint16_t) * (size_t)(e1 - e0));
                    if (!rc) rc = upload(h, s, colin_d[i] + e0, indices + s.nnz_begin + e0, sizeof(int) * (size_t)(e1 - e0));
                    if (rc) return rc;
                }
                CU(cudaEventRecord(s.ev_up, s.stream));
                CU(cudaStreamWaitEvent(s.aux, s.ev_up, 0));
                if (r1 > r0) {
                    k_col_signature<<<grid_for(r1 - r0, 256, s.n_sm * 16), 256, 0, s.aux>>>(
                        s.indptr + r0, r1 - r0, colin_d[i], raw_d[i], K, ((unsigned long long)s.world_rank << 40) + (unsigned long long)r0,
                        cnt_d[i], s.bad);
                    LAUNCH(h);
                    if (!permute) {
                        k_build_q<<<grid_for(e1 - e0, 256, s.n_sm * 16), 256, 0, s.aux>>>(raw_d[i] + e0, colin_d[i] + e0, lut_d[i], lut_len,
                                                                                       nullptr, s.q + e0, s.col + e0, e1 - e0, s.bad);
                        LAUNCH(h);
                        k_row_init<<<grid_for(r1 - r0, 256, s.n_sm * 16), 256, 0, s.aux>>>(Csr{s.indptr + r0, s.col, s.q, r1 - r0}, K, s.wy + r0,
                                                                                        s.scalars, s.pisum0);
                        LAUNCH(h);
                    }
                    CU(cudaGetLastError());
                }
                r0 = r1;
            }
            // the main stream goes on once the second one is through
            CU(cudaEventRecord(s.ev_up, s.aux));
            CU(cudaStreamWaitEvent(s.stream, s.ev_up, 0));
        }
        if (tm.on) { cudaStreamSynchronize(s.stream); }
        tm.lap("  entries H2D + signatures + Q");
    }
    if (tm.on) sync_all(h);
    tm.lap("column signatures");
    {   // global per-locus entry counts -> internal numbering (descending count, ties by original index)
        static_assert(sizeof(unsigned long long) == 8, "");
        for (int i = 0; i < n_local; ++i) h->shards[i].xchg_ptr = cnt_d[i];
        ALLREDUCE(h, s.xchg_ptr, (size_t)K * 5, ncclUint64, ncclSum);
        for (auto& s : h->shards) {
            int bad = 0;
            CU(cudaSetDevice(s.dev));
            CU(cudaMemcpyAsync(&bad, s.bad, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
            CU(cudaStreamSynchronize(s.stream));
            if (bad & 1) return fail(TSC_ERR_ARG, "column index out of range [0, n_cols)");
        }
        std::vector<unsigned long long> cnt((size_t)K * 5);
        Shard& s0 = h->shards[0];
        CU(cudaSetDevice(s0.dev));
        CU(cudaMemcpyAsync(cnt.data(), cnt_d[0], sizeof(unsigned long long) * K * 5, cudaMemcpyDeviceToHost, s0.stream));
        CU(cudaStreamSynchronize(s0.stream));
        h->d2h += sizeof(unsigned long long) * K * 5;
        h->pos_count.assign(cnt.begin() + 3 * (size_t)K, cnt.begin() + 4 * (size_t)K);      // entries with a positive score, per locus
        h->inv.resize(K);
        std::iota(h->inv.begin(), h->inv.end(), 0);
        if (cfg.permute_columns)
            std::stable_sort(h->inv.begin(), h->inv.end(), [&](int a, int b) { return cnt[a] > cnt[b]; });
        h->perm.resize(K);
        for (int i = 0; i < K; ++i) h->perm[h->inv[i]] = i;
        // classes of identical columns (same count and both signatures); empty loci are trivially identical too
        std::vector<int> order(K);
        std::iota(order.begin(), order.end(), 0);
        // a class = loci that agree in both exact counts and all three 64-bit signatures
        auto key_less = [&](int a, int b) {
            for (int w = 0; w < 5; ++w) {
                const unsigned long long x = cnt[(size_t)w * K + a], y = cnt[(size_t)w * K + b];
                if (x != y) return x < y;
            }
            return a < b;
        };
        auto key_eq = [&](int a, int b) {
            for (int w = 0; w < 5; ++w) if (cnt[(size_t)w * K + a] != cnt[(size_t)w * K + b]) return false;
            return true;
        };
        std::sort(order.begin(), order.end(), key_less);
        std::vector<int> rep_internal(K);
        h->n_dup_loci = 0;
        for (int i = 0; i < K;) {
            int j = i;
            while (j < K && key_eq(order[i], order[j])) ++j;
            for (int t = i; t < j; ++t) rep_internal[h->perm[order[t]]] = h->perm[order[i]];   // order[i] = smallest original index
            h->n_dup_loci += (j - i - 1);
            i = j;
        }
        if (h->n_dup_loci > 0) {
            for (auto& s : h->shards) {
                CU(cudaSetDevice(s.dev));
                CU(dev_malloc(&s.rep, sizeof(int) * K));
                CU(cudaMemcpyAsync(s.rep, rep_internal.data(), sizeof(int) * K, cudaMemcpyHostToDevice, s.stream));
                CU(cudaStreamSynchronize(s.stream));
            }
        }
    }
    tm.lap("locus classes");
    for (int i = 0; i < n_local; ++i) {
        Shard& s = h->shards[i];
        CU(cudaSetDevice(s.dev));
        // (the K-sized arrays live in the shard's small slab, zeroed when it was allocated)
        CU(cudaMemcpyAsync(s.perm, h->perm.data(), sizeof(int) * K, cudaMemcpyHostToDevice, s.stream));
        if (permute) {
            k_build_q<<<grid_for(s.nnz, 256, s.n_sm * 16), 256, 0, s.stream>>>(raw_d[i], colin_d[i], lut_d[i], lut_len, s.perm, s.q, s.col,
                                                                            s.nnz, s.bad);
            LAUNCH(h);
            CU(cudaGetLastError());
        }
        static_assert(sizeof(EmState) * 2 <= 256, "two loop states fit a pinned slot");
        if (!s.st_host) s.st_host = (EmState*)pinned_word_take();
        if (!s.st_host) return fail(TSC_ERR_ALLOC, "no page-locked host memory for the loop state");
        CU(cudaEventCreateWithFlags(&s.ev_poll[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&s.ev_poll[1], cudaEventDisableTiming));
        s.grid_rows = s.n_sm * 4;                   // 512-thread blocks, persistent grid-stride
        s.grid_tiles = s.n_sm * 2;                  // refined below from the occupancy of the tile kernel
        k_log_table<<<1, kLogTab, 0, s.stream>>>(s.log_tab);
        LAUNCH(h);
        if (permute) {
            k_row_init<<<grid_for(s.n_rows, 256, s.n_sm * 16), 256, 0, s.stream>>>(csr_of(s), K, s.wy, s.scalars, s.pisum0);
            LAUNCH(h);
        }
        CU(cudaGetLastError());
        k_fill<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.ones, K, 1.0);
        LAUNCH(h);
        k_fill<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.pi, K, 1.0 / K);        // model.py:667
        LAUNCH(h);
        k_fill<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.theta, K, 1.0 / K);     // model.py:673
        LAUNCH(h);
        k_mul<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.pi, s.theta, s.pt, K);
        LAUNCH(h);
        CU(cudaGetLastError());
    }
    if (tm.on) sync_all(h);
    tm.lap("build Q + row init");
    if (h->kernel == TSC_KERNEL_ELL) {
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            CU(cudaFuncSetAttribute(k_ell<ELL_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kEllSmem));
            CU(cudaFuncSetAttribute(k_ell<ELL_LNL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ell_smem_bytes<ELL_LNL>()));
            CU(cudaFuncSetAttribute(k_ell<ELL_REASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ell_smem_bytes<ELL_REASSIGN>()));
            CU(cudaFuncSetAttribute(k_ell_long<ELL_FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ell_long_smem_bytes<ELL_FUSED>()));
            CU(cudaFuncSetAttribute(k_ell_long<ELL_LNL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ell_long_smem_bytes<ELL_LNL>()));
            CU(cudaFuncSetAttribute(k_ell_long<ELL_REASSIGN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ell_long_smem_bytes<ELL_REASSIGN>()));
            Arena arena;        // the raw scores are dead once Q is built: their memory serves the clustering's temporaries
            CU(cudaStreamSynchronize(s.stream));
            if (s.raw && s.nnz > 0) { arena.base = (char*)s.raw; arena.cap = sizeof(uint16_t) * (size_t)s.nnz; }
            int rc = build_ell(h, s, tm, &arena);
            if (rc) return rc;
        }
        tm.lap("clustered ELL stream");
        // the stream's flushes are a few thousand coalesced REDs per iteration: replicas only pay for the flat tiles
        bool small_residual = cfg.replicas <= 0;
        for (auto& s : h->shards) small_residual = small_residual && s.res_amb_nnz * 20 < std::max<long long>(s.nnz, 1);
        if (small_residual) h->R_used = std::min(h->R, 2);
    }
    // totals over all shards (model.py:691-699)
    ALLREDUCE(h, s.scalars, 2, ncclFloat64, ncclSum);
    ALLREDUCE(h, s.scalars + 2, 1, ncclFloat64, ncclMax);
    ALLREDUCE(h, s.pisum0, (size_t)K, ncclFloat64, ncclSum);
    for (auto& s : h->shards) {
        int bad = 0;
        double part[3];
        CU(cudaSetDevice(s.dev));
        CU(cudaMemcpyAsync(&bad, s.bad, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
        CU(cudaMemcpyAsync(part, s.scalars, sizeof(double) * 3, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        if (bad & 2) return fail(TSC_ERR_ARG, "raw score >= lut_len");
        Consts c;
        c.total_wt = part[0];
        c.ambig_wt = part[1];
        c.wmax = part[2];
        c.pi_prior_wt = pi_prior * c.wmax;                           // model.py:696
        c.theta_prior_wt = theta_prior * c.wmax;                     // model.py:697
        c.theta_denom = c.ambig_wt + c.theta_prior_wt * K;           // model.py:734
        c.pi_denom = c.total_wt + c.pi_prior_wt * K;                 // model.py:739
        h->consts = c;
        CU(cudaMemcpyAsync(s.consts, &c, sizeof(Consts), cudaMemcpyHostToDevice, s.stream));
        CU(cudaStreamSynchronize(s.stream));
    }
    // shared-memory staging of the pi*theta table (optional)
    {
        int want = cfg.smem_table_cols;
        if (want < 0) want = 0;                       // auto: gather through L1/L2 (measured faster when loci are skewed)
        want = std::min(want, K);
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            int max_optin = 0;
            CU(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, s.dev));
            const size_t scratch = sizeof(double) * kTileWarps * kScratch;
            const int fit = (int)((max_optin - scratch - 1024) / sizeof(double));
            s.s_cols = std::min(want, std::max(fit, 0));
            s.smem_tiles = scratch + sizeof(double) * s.s_cols;
            const int big = (int)std::max(s.smem_tiles, scratch);
            CU(cudaFuncSetAttribute(k_tiles<TILE_FUSED, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CU(cudaFuncSetAttribute(k_tiles<TILE_FUSED, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
            CU(cudaFuncSetAttribute(k_tiles<TILE_FUSED, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            CU(cudaFuncSetAttribute(k_tiles<TILE_FUSED, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            CU(cudaFuncSetAttribute(k_tiles<TILE_Z, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            CU(cudaFuncSetAttribute(k_tiles<TILE_Z, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            CU(cudaFuncSetAttribute(k_tiles<TILE_LNL, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            CU(cudaFuncSetAttribute(k_tiles<TILE_LNL, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)scratch));
            if (s.s_cols > 0) {
                h->smem_tab = true;
                s.grid_tiles = s.n_sm;
            } else {
                int per_sm = 0;
                CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_tiles<TILE_FUSED, false, true>, kTileThreads, scratch));
                s.grid_tiles = s.n_sm * std::max(per_sm, 1);
            }
        }
    }
    int rc_final = sync_all(h);
    tm.lap("constants + finish");
    cleanup_tmp();
    tm.lap("free temporaries");
    return rc_final;
}

extern "C" int tsc_create(tsc_handle** out, const tsc_config* cfg_in, int64_t n_rows, int32_t n_cols, int64_t nnz,
                          const void* indptr, int32_t indptr_bytes, const int32_t* indices, const uint16_t* raw,
                          const double* q_lut, int32_t lut_len, double pi_prior, double theta_prior) {
    if (!out) return fail(TSC_ERR_ARG, "out is NULL");
    *out = nullptr;
    if (n_rows < 0 || n_cols <= 0 || nnz < 0) return fail(TSC_ERR_ARG, "negative or empty dimensions");
    if (!indptr || (nnz > 0 && (!indices || !raw))) return fail(TSC_ERR_ARG, "NULL CSR array");
    if (indptr_bytes != 4 && indptr_bytes != 8) return fail(TSC_ERR_ARG, "indptr_bytes must be 4 or 8");
    if (!q_lut || lut_len <= 0) return fail(TSC_ERR_ARG, "q_lut missing");
    tsc_config cfg;
    if (cfg_in) cfg = *cfg_in; else tsc_config_default(&cfg);
    tsc_handle* h = new (std::nothrow) tsc_handle();
    if (!h) return fail(TSC_ERR_ALLOC, "out of host memory");
    const CreateInput in{n_rows, n_cols, nnz, indptr, indptr_bytes, indices, raw, q_lut, lut_len, pi_prior, theta_prior};
    int rc = create_impl(h, cfg, in);
    if (rc) { std::string keep = g_err; tsc_destroy(h); g_err = keep; return rc; }
    *out = h;
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- getters
extern "C" int tsc_get_constants(tsc_handle* h, double* scalars5, double* pisum0) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    if (scalars5) {
        scalars5[0] = h->consts.total_wt; scalars5[1] = h->consts.ambig_wt; scalars5[2] = h->consts.wmax;
        scalars5[3] = h->consts.pi_prior_wt; scalars5[4] = h->consts.theta_prior_wt;
    }
    if (pisum0) { Shard& s = h->shards[0]; CU(cudaSetDevice(s.dev)); return get_kvec(h, s, s.pisum0, pisum0); }
    return TSC_OK;
}

extern "C" int tsc_get_row_info(tsc_handle* h, uint8_t* y_rows, double* w_rows) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    std::vector<uint8_t> y;
    std::vector<double> w;
    const bool compact = !h->rowmap.empty() || h->n_rows != h->n_rows_user;
    if (compact) { if (y_rows) y.resize(h->n_rows); if (w_rows) w.resize(h->n_rows); }
    uint8_t* yh = compact ? y.data() : y_rows;
    double* wh = compact ? w.data() : w_rows;
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        uint8_t* yd = nullptr; double* wd = nullptr;
        if (y_rows) CU(dev_malloc(&yd, std::max<long long>(s.n_rows, 1)));
        if (w_rows) CU(dev_malloc(&wd, sizeof(double) * std::max<long long>(s.n_rows, 1)));
        k_row_info<<<grid_for(s.n_rows, 256, s.n_sm * 16), 256, 0, s.stream>>>(csr_of(s), s.wy, yd, wd);
        LAUNCH(h);
        if (y_rows) CU(cudaMemcpyAsync(yh + s.row_begin, yd, s.n_rows, cudaMemcpyDeviceToHost, s.stream));
        if (w_rows) CU(cudaMemcpyAsync(wh + s.row_begin, wd, sizeof(double) * s.n_rows, cudaMemcpyDeviceToHost, s.stream));
        CU(cudaStreamSynchronize(s.stream));
        if (yd) dev_free(yd);
        if (wd) dev_free(wd);
        h->d2h += (y_rows ? s.n_rows : 0) + (w_rows ? 8 * s.n_rows : 0);
    }
    if (compact) {
        if (y_rows) { memset(y_rows, 0, h->n_rows_user); for (long long r = 0; r < h->n_rows; ++r) y_rows[h->rowmap[r]] = y[r]; }
        if (w_rows) { std::fill(w_rows, w_rows + h->n_rows_user, 0.0); for (long long r = 0; r < h->n_rows; ++r) w_rows[h->rowmap[r]] = w[r]; }
    }
    return TSC_OK;
}

extern "C" int tsc_get_q(tsc_handle* h, double* q_data) {
    if (!h || !q_data) return fail(TSC_ERR_ARG, "NULL argument");
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        CU(cudaMemcpyAsync(q_data + s.nnz_begin, s.q, sizeof(double) * s.nnz, cudaMemcpyDeviceToHost, s.stream));
        h->d2h += sizeof(double) * s.nnz;
    }
    return sync_all(h);
}

extern "C" int tsc_get_params(tsc_handle* h, double* pi, double* theta, double* pi_init, double* theta_init) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    Shard& s = h->shards[0];
    CU(cudaSetDevice(s.dev));
    int rc;
    if (pi && (rc = get_kvec(h, s, s.pi, pi))) return rc;
    if (theta && (rc = get_kvec(h, s, s.theta, theta))) return rc;
    if ((pi_init || theta_init) && h->n_iter == 0) return fail(TSC_ERR_STATE, "pi_init/theta_init exist only after em()");
    if (pi_init && (rc = get_kvec(h, s, s.pi_init, pi_init))) return rc;
    if (theta_init && (rc = get_kvec(h, s, s.theta_init, theta_init))) return rc;
    return TSC_OK;
}

extern "C" int tsc_set_params(tsc_handle* h, const double* pi, const double* theta) {
    if (!h || !pi || !theta) return fail(TSC_ERR_ARG, "NULL argument");
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        int rc;
        if ((rc = put_kvec(h, s, pi, s.pi))) return rc;
        if ((rc = put_kvec(h, s, theta, s.theta))) return rc;
        k_mul<<<grid_for(h->K, 256, 1 << 20), 256, 0, s.stream>>>(s.pi, s.theta, s.pt, h->K);
        LAUNCH(h);
        CU(cudaGetLastError());
    }
    return sync_all(h);
}

extern "C" int tsc_get_counters(tsc_handle* h, int64_t* launches, int64_t* h2d_bytes, int64_t* d2h_bytes) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    if (launches) *launches = h->launches;
    if (h2d_bytes) *h2d_bytes = h->h2d;
    if (d2h_bytes) *d2h_bytes = h->d2h;
    return TSC_OK;
}

extern "C" int tsc_get_layout_stats(tsc_handle* h, int64_t* out8) {
    if (!h || !out8) return fail(TSC_ERR_ARG, "NULL argument");
    const Shard& s = h->shards[0];
    out8[0] = s.ell_bytes; out8[1] = s.ell_slices; out8[2] = s.ell_reads; out8[3] = s.ell_entries;
    out8[4] = s.res_amb_rows; out8[5] = s.res_amb_nnz; out8[6] = s.ell_grid; out8[7] = s.ell_long;
    return TSC_OK;
}

extern "C" int tsc_get_transport(tsc_handle* h, int32_t* transport_out) {
    if (!h || !transport_out) return fail(TSC_ERR_ARG, "NULL argument");
    *transport_out = h->transport;
    return TSC_OK;
}

extern "C" int tsc_get_em_device_ms(tsc_handle* h, float* ms_out) {
    if (!h || !ms_out) return fail(TSC_ERR_ARG, "NULL argument");
    *ms_out = h->em_ms;
    return TSC_OK;
}

extern "C" void* tsc_pinned_alloc(uint64_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable);
    if (e != cudaSuccess) { fail(TSC_ERR_ALLOC, std::string("cudaHostAlloc: ") + cudaGetErrorString(e)); return nullptr; }
    return p;
}
extern "C" void tsc_pinned_free(void* p) { if (p) cudaFreeHost(p); }

extern "C" const char* tsc_create_laps(tsc_handle* h) { return h ? h->create_laps.c_str() : ""; }

extern "C" int tsc_allreduce_f64(tsc_handle* h, double* inout, int32_t n, int32_t op) {
    if (!h || !inout || n <= 0 || n > 8 || (op != 0 && op != 1)) return fail(TSC_ERR_ARG, "bad argument");
    if (h->world == 1) return TSC_OK;
    // every local shard contributes the same host values; divide sums by the local shard count afterwards
    const int n_local = (int)h->shards.size();
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        CU(cudaMemcpyAsync(s.scalars, inout, sizeof(double) * n, cudaMemcpyHostToDevice, s.stream));
    }
    if (op == 0) ALLREDUCE(h, s.scalars, (size_t)n, ncclFloat64, ncclSum);
    else ALLREDUCE(h, s.scalars, (size_t)n, ncclFloat64, ncclMax);
    Shard& s0 = h->shards[0];
    CU(cudaSetDevice(s0.dev));
    CU(cudaMemcpyAsync(inout, s0.scalars, sizeof(double) * n, cudaMemcpyDeviceToHost, s0.stream));
    int rc = sync_all(h);
    if (rc) return rc;
    if (op == 0 && n_local > 1) for (int i = 0; i < n; ++i) inout[i] /= n_local;
    return TSC_OK;
}

extern "C" int tsc_time_pass(tsc_handle* h, int32_t pass_id, int32_t reps, float* mean_ms) {
    if (!h || !mean_ms || reps <= 0) return fail(TSC_ERR_ARG, "bad argument");
    Shard& s = h->shards[0];
    CU(cudaSetDevice(s.dev));
    double* zd = nullptr;
    if (pass_id == 1) { int rc = ensure_tiles(h, s); if (rc) return rc; }
    if (pass_id == 1) CU(dev_malloc(&zd, sizeof(double) * std::max<long long>(s.nnz, 1)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    const double* ta = h->em_done ? s.pt_prev : s.pt;
    const double* tu = h->em_done ? s.pi_prev : s.pi;
    cudaError_t err = cudaSuccess;
    for (int r = -1; r < reps && err == cudaSuccess; ++r) {      // r = -1 is a warm-up
        if (r == 0) cudaEventRecord(e0, s.stream);
        TileArgs a{};
        a.tiles = s.tiles; a.n_tiles = s.n_tiles; a.q = s.q; a.col = s.col; a.K = h->K; a.R = h->R;
        switch (pass_id) {
            case 0: {
                const int rc0 = launch_fused(h, s, false);
                if (rc0) err = cudaErrorUnknown;
                h->launches -= 1;            // counted below
            } break;
            case 1:
                a.tab_amb = ta; a.tab_uni = tu; a.z_out = zd;
                launch_tiles<TILE_Z>(s, a, false);
                break;
            case 2: {
                int np = 0;
                const long long l0 = h->launches;
                if (launch_lnl_kernels(h, s, nullptr, ta, tu, s.pt, s.pi, &np)) err = cudaErrorUnknown;
                h->launches = l0;             // counted below
            } break;
            case 3: {
                const long long l0 = h->launches;
                if (launch_reassign_sums(h, s, TSC_EXCLUDE, 0.9, ta, tu, s.colsum, nullptr)) err = cudaErrorUnknown;
                h->launches = l0;             // counted below
            } break;
            default:
                err = cudaErrorInvalidValue;
        }
        LAUNCH(h);
        if (err == cudaSuccess) err = cudaGetLastError();
    }
    if (err == cudaSuccess) err = cudaEventRecord(e1, s.stream);
    if (err == cudaSuccess) err = cudaEventSynchronize(e1);
    float ms = 0.f;
    if (err == cudaSuccess) err = cudaEventElapsedTime(&ms, e0, e1);
    if (pass_id == 0) cudaMemsetAsync(s.acc, 0, sizeof(double) * h->K * h->R, s.stream);   // discard the sums
    cudaStreamSynchronize(s.stream);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (zd) dev_free(zd);
    if (err != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("tsc_time_pass: ") + cudaGetErrorString(err));
    *mean_ms = ms / reps;
    return TSC_OK;
}

#ifdef TSC_ELL_TRACE
// debug build only: start / end (globaltimer ns) of every CTA of the last k_ell<ELL_FUSED> launch on shard 0
extern "C" int tsc_debug_ell_trace(tsc_handle* h, uint64_t* out, int32_t n_ctas) {
    if (!h || !out) return fail(TSC_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(h->shards[0].dev));
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpyFromSymbol(out, g_ell_trace, sizeof(uint64_t) * 2 * std::min(n_ctas, 8192)));
    return TSC_OK;
}
#endif

extern "C" int tsc_get_kernel_times(tsc_handle* h, float* ms_out, int32_t max_n, int32_t* n_out) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    const int n = (int)std::min<size_t>(h->kernel_ms.size(), (size_t)std::max(max_n, 0));
    if (ms_out) for (int i = 0; i < n; ++i) ms_out[i] = h->kernel_ms[i];
    if (n_out) *n_out = (int)h->kernel_ms.size();
    return TSC_OK;
}

extern "C" int tsc_get_tail_times(tsc_handle* h, float* ms_out, int32_t max_n, int32_t* n_out) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    const int n = (int)std::min<size_t>(h->tail_ms.size(), (size_t)std::max(max_n, 0));
    if (ms_out) for (int i = 0; i < n; ++i) ms_out[i] = h->tail_ms[i];
    if (n_out) *n_out = (int)h->tail_ms.size();
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- EM
static int launch_fused(tsc_handle* h, Shard& s, bool gated) {
    const EmState* st = gated ? s.st : nullptr;
    if (h->kernel == TSC_KERNEL_ROWS) {
        launch_rows(h->G, [&](auto g) {
            k_fused_rows<decltype(g)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), s.wy, s.pt, s.acc, h->K, h->R_used, st);
        });
        LAUNCH(h);
    } else if (h->kernel == TSC_KERNEL_ELL) {
        // the clustered stream, then whatever does not fit a slice through the flat tiles of the residual CSR
        if (s.ell_slices > 0) {
            const bool measure = gated && s.ell_rebal_left > 0;
            EllArgs e{s.ell_stream, s.ell_index, s.ell_range, s.ell_slices, s.pt, s.acc, h->K, h->R_used, st, nullptr, nullptr, nullptr, 0, 0.0, nullptr, nullptr,
                      measure ? s.ell_cta_ns : nullptr};
            k_ell<ELL_FUSED><<<s.ell_grid, 32, kEllSmem, s.stream>>>(e);
            LAUNCH(h);
            if (measure) {
                // (the long-read records follow the slices in the stream: the slices end where they begin)
                const unsigned end16 = (unsigned)((s.ell_long > 0 ? s.ell_slice_bytes : s.ell_bytes) >> 4);
                k_ell_rebalance<<<1, 1024, 0, s.stream>>>(s.ell_index, s.ell_slices, end16, s.ell_grid, s.ell_range, s.ell_cta_ns, s.ell_rebal, st);
                LAUNCH(h);
                --s.ell_rebal_left;
            }
        }
        if (s.ell_long > 0) {
            EllArgs e{s.ell_stream, s.ell_index + s.ell_slices, s.ell_lrange, s.ell_long, s.pt, s.acc, h->K, h->R_used, st, nullptr, nullptr, nullptr, 0, 0.0, nullptr, nullptr};
            k_ell_long<ELL_FUSED><<<s.ell_lgrid, 32, ell_long_smem_bytes<ELL_FUSED>(), s.stream>>>(e);
            LAUNCH(h);
        }
        if (s.res_amb_rows > 0) {         // (a residual of unique reads only has nothing to add)
            TileArgs a{};
            a.tiles = s.res_tiles_amb; a.n_tiles = s.res_amb_tiles; a.q = s.res_q; a.col = s.res_col; a.wy = s.res_wy; a.tab_amb = s.pt;
            a.acc = s.acc; a.K = h->K; a.R = h->R_used; a.s_cols = 0; a.st = st;
            launch_tiles<TILE_FUSED>(s, a, false, s.res_amb_long * 200 > s.res_amb_tiles ? 1 : 0);
            LAUNCH(h);
        }
    } else {
        TileArgs a{};
        a.tiles = s.tiles; a.n_tiles = s.n_tiles; a.q = s.q; a.col = s.col; a.wy = s.wy; a.tab_amb = s.pt;
        a.acc = s.acc; a.K = h->K; a.R = h->R_used; a.s_cols = s.s_cols; a.st = st;
        launch_tiles<TILE_FUSED>(s, a, s.s_cols > 0);
        LAUNCH(h);
    }
    CU(cudaGetLastError());
    return TSC_OK;
}

// global log-likelihood into s.scalars[4] of every shard (model.py:744-760)
// One shard's log-likelihood with z regenerated from the E-step tables (ta, tu) and `inner` inside log1p: per-CTA
// partials into s.partials, returns how many.  Clustered-stream layout: the stream kernel for its reads, the flat
// tiles of the residual CSR for everybody else; otherwise the flat tiles of the whole shard.
static int launch_lnl_kernels(tsc_handle* h, Shard& s, const EmState* st, const double* ta, const double* tu,
                              const double* ia, const double* iu, int* nparts_out) {
    int nparts = 0;
    TileArgs a{};
    a.tab_amb = ta; a.tab_uni = tu; a.inner_amb = ia; a.inner_uni = iu; a.K = h->K; a.st = st; a.log_tab = s.log_tab;
    if (h->kernel == TSC_KERNEL_ELL) {
        if (s.ell_slices > 0) {
            EllArgs e{s.ell_stream, s.ell_index, s.ell_range_lnl, s.ell_slices, ta, nullptr, h->K, 1, st, ia, s.partials, s.log_tab, 0, 0.0, nullptr, nullptr};
            k_ell<ELL_LNL><<<s.ell_grid_lnl, 32, ell_smem_bytes<ELL_LNL>(), s.stream>>>(e);
            LAUNCH(h);
            nparts = s.ell_grid_lnl;
        }
        if (s.ell_long > 0) {
            EllArgs e{s.ell_stream, s.ell_index + s.ell_slices, s.ell_lrange_lnl, s.ell_long, ta, nullptr, h->K, 1, st, ia, s.partials + nparts,
                      s.log_tab, 0, 0.0, nullptr, nullptr};
            k_ell_long<ELL_LNL><<<s.ell_lgrid_lnl, 32, ell_long_smem_bytes<ELL_LNL>(), s.stream>>>(e);
            LAUNCH(h);
            nparts += s.ell_lgrid_lnl;
        }
        if (s.res_rows > 0) {
            a.tiles = s.res_tiles; a.n_tiles = s.res_n_tiles; a.q = s.res_q; a.col = s.res_col; a.partials = s.partials + nparts;
            launch_tiles<TILE_LNL>(s, a, false, s.res_n_long * 200 > s.res_n_tiles ? 1 : 0);
            LAUNCH(h);
            nparts += tiles_grid(s, a.n_tiles);
        }
    } else {
        a.tiles = s.tiles; a.n_tiles = s.n_tiles; a.q = s.q; a.col = s.col; a.partials = s.partials;
        launch_tiles<TILE_LNL>(s, a, false);
        LAUNCH(h);
        nparts = tiles_grid(s, a.n_tiles);
    }
    CU(cudaGetLastError());
    *nparts_out = nparts;
    return TSC_OK;
}

// global log-likelihood into s.scalars[4] of every shard (model.py:744-760)
static int launch_lnl(tsc_handle* h, const double* (*zin_of)(Shard&), bool from_prev, bool gated,
                      const double* (*ia)(Shard&), const double* (*iu)(Shard&)) {
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        const double* zin = zin_of ? zin_of(s) : nullptr;
        const EmState* st = gated ? s.st : nullptr;
        int nparts = s.grid_rows;
        if (!zin && h->kernel != TSC_KERNEL_ROWS) {
            int rc = launch_lnl_kernels(h, s, st, s.pt_prev, s.pi_prev, ia(s), iu(s), &nparts);
            if (rc) return rc;
        } else {
            launch_rows(h->G, [&](auto g) {
                k_lnl_rows<decltype(g)::value><<<s.grid_rows, 512, 0, s.stream>>>(
                    csr_of(s), zin, from_prev ? s.pt_prev : nullptr, from_prev ? s.pi_prev : nullptr, ia(s), iu(s), s.partials, st);
            });
            LAUNCH(h);
            CU(cudaGetLastError());
        }
        // (when the loop is already done the partials are stale; k_lnl_control ignores the value)
        k_sum_partials<<<1, 1024, 0, s.stream>>>(s.partials, nparts, s.scalars + 4);
        LAUNCH(h);
        CU(cudaGetLastError());
    }
    ALLREDUCE(h, s.scalars + 4, 1, ncclFloat64, ncclSum);
    return TSC_OK;
}

extern "C" int tsc_em(tsc_handle* h, int32_t max_iter, double eps, int32_t use_likelihood, double* diffs_out,
                      double* lnls_out, int32_t* n_iter, int32_t* converged, double* final_lnl) {
    if (!h) return fail(TSC_ERR_ARG, "handle is NULL");
    if (use_likelihood && !lnls_out) return fail(TSC_ERR_ARG, "lnls_out required with use_likelihood");
    { int rc = peer_ready(h); if (rc) return rc; }
    const int T = std::max(1, (int)max_iter);   // the reference's loop body always runs once (model.py:771-794)
    const int K = h->K;
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        if (s.diffs_cap < T) {
            if (s.diffs) dev_free(s.diffs);
            if (s.lnls) dev_free(s.lnls);
            CU(dev_malloc(&s.diffs, sizeof(double) * T));
            CU(dev_malloc(&s.lnls, sizeof(double) * T));
            s.diffs_cap = T;
        }
        EmState st0{};
        st0.lnl_prev = h->lnl;          // self.lnl carries over between em() calls (model.py:683,786)
        st0.lnl = h->lnl;
        s.st_host[0] = st0;
        s.st_host[1] = st0;
        CU(cudaMemcpyAsync(s.st, &s.st_host[0], sizeof(EmState), cudaMemcpyHostToDevice, s.stream));
    }
    Shard& s0 = h->shards[0];
    CU(cudaSetDevice(s0.dev));
    // kernel-time events: a bounded pool (the first kTimedIters iterations of a call are timed), created on demand
    constexpr int kTimedIters = 512;
    const int n_timed = std::min(T, kTimedIters);
    while ((int)s0.ev_k.size() < 3 * n_timed) {
        cudaEvent_t e;
        CU(cudaEventCreate(&e));
        s0.ev_k.push_back(e);
    }
    if (!s0.ev_em[0]) { CU(cudaEventCreate(&s0.ev_em[0])); CU(cudaEventCreate(&s0.ev_em[1])); }
    CU(cudaEventRecord(s0.ev_em[0], s0.stream));
    const int poll_every = 4;
    int issued = 0, polls = 0;
    bool stop = false;
    for (int it = 0; it < T && !stop; ++it) {
        for (auto& s : h->shards) {
            CU(cudaSetDevice(s.dev));
            if (&s == &s0 && it < n_timed) CU(cudaEventRecord(s0.ev_k[3 * it], s.stream));
            int rc = launch_fused(h, s, true);
            if (rc) return rc;
            if (&s == &s0 && it < n_timed) CU(cudaEventRecord(s0.ev_k[3 * it + 1], s.stream));
            if (h->transport == TSC_TRANSPORT_NCCL) {
                k_reduce_replicas<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.acc, K, h->R_used, s.thetasum, s.st);
                LAUNCH(h);
                CU(cudaGetLastError());
            }
        }
        if (h->transport == TSC_TRANSPORT_NCCL) {
            ALLREDUCE(h, s.thetasum, (size_t)K, ncclFloat64, ncclSum);
            for (auto& s : h->shards) {
                CU(cudaSetDevice(s.dev));
                UpdateArgs a{s.thetasum, s.pisum0, s.consts, s.pi, s.theta, s.pt, s.pi_prev, s.theta_prev, s.pt_prev,
                             s.pi_init, s.theta_init, s.st, s.diffs, K, T, use_likelihood ? 1 : 0, eps, s.rep};
                k_update<<<1, 1024, 0, s.stream>>>(a);
                LAUNCH(h);
                CU(cudaGetLastError());
            }
        } else {
            // replica sum + exchange over peer memory + update + loop control: one kernel (tsc_peer.cuh)
            for (auto& s : h->shards) {
                CU(cudaSetDevice(s.dev));
                TailArgs a{UpdateArgs{nullptr, s.pisum0, s.consts, s.pi, s.theta, s.pt, s.pi_prev, s.theta_prev, s.pt_prev,
                                      s.pi_init, s.theta_init, s.st, s.diffs, K, T, use_likelihood ? 1 : 0, eps, s.rep},
                           peer_args(h, s), s.acc, h->R_used, h->epoch_iter, s.tail_partials, s.tail_ticket};
                k_tail<<<h->nb_tail, kTailThreads, 0, s.stream>>>(a);
                LAUNCH(h);
                CU(cudaGetLastError());
            }
        }
        if (it < n_timed) { CU(cudaSetDevice(s0.dev)); CU(cudaEventRecord(s0.ev_k[3 * it + 2], s0.stream)); }
        if (use_likelihood) {
            int rc = launch_lnl(h, nullptr, true, true,
                                [](Shard& s) -> const double* { return s.pt; }, [](Shard& s) -> const double* { return s.pi; });
            if (rc) return rc;
            for (auto& s : h->shards) {
                CU(cudaSetDevice(s.dev));
                k_lnl_control<<<1, 1, 0, s.stream>>>(s.st, s.scalars + 4, s.lnls, eps, T);
                LAUNCH(h);
                CU(cudaGetLastError());
            }
        }
        ++issued;
        if (issued % poll_every == 0 || it == T - 1) {
            CU(cudaSetDevice(s0.dev));
            const int slot = polls & 1;
            CU(cudaMemcpyAsync(&s0.st_host[slot], s0.st, sizeof(EmState), cudaMemcpyDeviceToHost, s0.stream));
            CU(cudaEventRecord(s0.ev_poll[slot], s0.stream));
            if (polls > 0) {
                CU(cudaEventSynchronize(s0.ev_poll[slot ^ 1]));
                if (s0.st_host[slot ^ 1].done) stop = true;
            }
            ++polls;
        }
    }
    int rc = sync_all(h);
    if (rc) return rc;
    EmState fin;
    CU(cudaSetDevice(s0.dev));
    CU(cudaMemcpy(&fin, s0.st, sizeof(EmState), cudaMemcpyDeviceToHost));
    h->epoch_iter += (unsigned long long)T;      // the same on every rank, whatever each of them launched or skipped
    if (fin.pad) return fail(TSC_ERR_STATE, "peer exchange timed out: another rank of this model stopped taking part");
    h->n_iter = fin.iter;
    h->converged = fin.converged;
    h->em_done = true;
    h->kernel_ms.clear();
    h->tail_ms.clear();
    for (int it = 0; it < std::min(std::min(issued, fin.iter), n_timed); ++it) {
        float ms = 0;
        CU(cudaEventElapsedTime(&ms, s0.ev_k[3 * it], s0.ev_k[3 * it + 1]));
        h->kernel_ms.push_back(ms);
        CU(cudaEventElapsedTime(&ms, s0.ev_k[3 * it + 1], s0.ev_k[3 * it + 2]));
        h->tail_ms.push_back(ms);
    }
    if (use_likelihood) {
        h->lnl = fin.lnl;
    } else {   // model.py:800-801
        rc = launch_lnl(h, nullptr, true, false,
                        [](Shard& s) -> const double* { return s.pt; }, [](Shard& s) -> const double* { return s.pi; });
        if (rc) return rc;
        CU(cudaSetDevice(s0.dev));
        CU(cudaMemcpyAsync(&h->lnl, s0.scalars + 4, sizeof(double), cudaMemcpyDeviceToHost, s0.stream));
        CU(cudaStreamSynchronize(s0.stream));
        rc = sync_all(h);
        if (rc) return rc;
    }
    CU(cudaSetDevice(s0.dev));
    CU(cudaEventRecord(s0.ev_em[1], s0.stream));
    CU(cudaEventSynchronize(s0.ev_em[1]));
    CU(cudaEventElapsedTime(&h->em_ms, s0.ev_em[0], s0.ev_em[1]));
    if (diffs_out) CU(cudaMemcpy(diffs_out, s0.diffs, sizeof(double) * fin.iter, cudaMemcpyDeviceToHost));
    if (lnls_out && use_likelihood) CU(cudaMemcpy(lnls_out, s0.lnls, sizeof(double) * fin.iter, cudaMemcpyDeviceToHost));
    h->d2h += sizeof(double) * fin.iter * (use_likelihood ? 2 : 1) + sizeof(EmState);
    if (n_iter) *n_iter = fin.iter;
    if (converged) *converged = fin.converged;
    if (final_lnl) *final_lnl = h->lnl;
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- estep / mstep / lnl
static int alloc_entries(Shard& s, double** p) {
    CU(cudaSetDevice(s.dev));
    CU(dev_malloc(p, sizeof(double) * std::max<long long>(s.nnz, 1)));
    return TSC_OK;
}

static int z_to_host(tsc_handle* h, int which, double* z_data) {
    // which: 0 = tables tmp_c/tmp_a (explicit pi, theta), 1 = *_prev (stored posterior), 2 = ones (Q.norm(1))
    for (auto& s : h->shards) {
        double* zd = nullptr;
        int rc = alloc_entries(s, &zd);
        if (rc) return rc;
        const double* ta = which == 0 ? s.tmp_c : which == 1 ? s.pt_prev : s.ones;
        const double* tu = which == 0 ? s.tmp_a : which == 1 ? s.pi_prev : s.ones;
        if (h->kernel != TSC_KERNEL_ROWS) {
            if ((rc = ensure_tiles(h, s))) { dev_free(zd); return rc; }
            TileArgs a{};
            a.tiles = s.tiles; a.n_tiles = s.n_tiles; a.q = s.q; a.col = s.col; a.tab_amb = ta; a.tab_uni = tu;
            a.K = h->K; a.z_out = zd;
            launch_tiles<TILE_Z>(s, a, false);
        } else {
            launch_rows(h->G, [&](auto g) {
                k_estep_rows<decltype(g)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), ta, tu, zd);
            });
        }
        LAUNCH(h);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(z_data + s.nnz_begin, zd, sizeof(double) * s.nnz, cudaMemcpyDeviceToHost, s.stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        dev_free(zd);
        if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("posterior export: ") + cudaGetErrorString(e));
        h->d2h += sizeof(double) * s.nnz;
    }
    return TSC_OK;
}

static int upload_pi_theta(tsc_handle* h, const double* pi, const double* theta) {
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        int rc;
        if ((rc = put_kvec(h, s, pi, s.tmp_a))) return rc;
        if ((rc = put_kvec(h, s, theta, s.tmp_b))) return rc;
        k_mul<<<grid_for(h->K, 256, 1 << 20), 256, 0, s.stream>>>(s.tmp_a, s.tmp_b, s.tmp_c, h->K);
        LAUNCH(h);
        CU(cudaGetLastError());
    }
    return TSC_OK;
}

extern "C" int tsc_estep(tsc_handle* h, const double* pi, const double* theta, double* z_data) {
    if (!h || !pi || !theta || !z_data) return fail(TSC_ERR_ARG, "NULL argument");
    int rc = upload_pi_theta(h, pi, theta);
    if (rc) return rc;
    return z_to_host(h, 0, z_data);
}

extern "C" int tsc_get_z(tsc_handle* h, int32_t initial, double* z_data) {
    if (!h || !z_data) return fail(TSC_ERR_ARG, "NULL argument");
    if (!initial && !h->em_done) return fail(TSC_ERR_STATE, "z exists only after em() (model.py:659,795)");
    return z_to_host(h, initial ? 2 : 1, z_data);
}

__global__ void k_mstep_tail(const double* __restrict__ thetasum, const double* __restrict__ pisum0, const Consts* __restrict__ c,
                             double* __restrict__ pi_hat, double* __restrict__ theta_hat, int K, const int* __restrict__ rep) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= K) return;
    const double ts = thetasum[rep ? rep[j] : j];
    theta_hat[j] = (ts + c->theta_prior_wt) / c->theta_denom;
    const double pisum = pisum0[j] + ts;
    pi_hat[j] = (pisum + c->pi_prior_wt) / c->pi_denom;
}

extern "C" int tsc_mstep(tsc_handle* h, const double* z_data, double* pi_hat, double* theta_hat) {
    if (!h || !z_data || !pi_hat || !theta_hat) return fail(TSC_ERR_ARG, "NULL argument");
    const int K = h->K;
    for (auto& s : h->shards) {
        double* zd = nullptr;
        int rc = alloc_entries(s, &zd);
        if (rc) return rc;
        cudaError_t e = cudaMemcpyAsync(zd, z_data + s.nnz_begin, sizeof(double) * s.nnz, cudaMemcpyHostToDevice, s.stream);
        h->h2d += sizeof(double) * s.nnz;
        if (e == cudaSuccess) {
            launch_rows(h->G, [&](auto g) {
                k_mstep_rows<decltype(g)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), s.wy, zd, s.acc, K, h->R);
            });
            LAUNCH(h);
            k_reduce_replicas<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.acc, K, h->R, s.tmp_c, nullptr);
            LAUNCH(h);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        dev_free(zd);
        if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("mstep: ") + cudaGetErrorString(e));
    }
    ALLREDUCE(h, s.tmp_c, (size_t)K, ncclFloat64, ncclSum);
    Shard& s = h->shards[0];
    CU(cudaSetDevice(s.dev));
    k_mstep_tail<<<grid_for(K, 256, 1 << 20), 256, 0, s.stream>>>(s.tmp_c, s.pisum0, s.consts, s.tmp_a, s.tmp_b, K, s.rep);
    LAUNCH(h);
    CU(cudaGetLastError());
    int rc = sync_all(h);
    if (rc) return rc;
    if ((rc = get_kvec(h, s, s.tmp_a, pi_hat))) return rc;
    return get_kvec(h, s, s.tmp_b, theta_hat);
}

extern "C" int tsc_calculate_lnl(tsc_handle* h, const double* z_data, const double* pi, const double* theta, double* lnl) {
    if (!h || !z_data || !pi || !theta || !lnl) return fail(TSC_ERR_ARG, "NULL argument");
    int rc = upload_pi_theta(h, pi, theta);
    if (rc) return rc;
    std::vector<double*> zd(h->shards.size(), nullptr);
    auto free_all = [&]() { for (size_t i = 0; i < zd.size(); ++i) if (zd[i]) { cudaSetDevice(h->shards[i].dev); dev_free(zd[i]); } };
    for (size_t i = 0; i < h->shards.size(); ++i) {
        Shard& s = h->shards[i];
        if ((rc = alloc_entries(s, &zd[i]))) { free_all(); return rc; }
        cudaError_t e = cudaMemcpyAsync(zd[i], z_data + s.nnz_begin, sizeof(double) * s.nnz, cudaMemcpyHostToDevice, s.stream);
        h->h2d += sizeof(double) * s.nnz;
        if (e == cudaSuccess) {
            launch_rows(h->G, [&](auto g) {
                k_lnl_rows<decltype(g)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), zd[i], nullptr, nullptr, s.tmp_c, s.tmp_a,
                                                                                s.partials, nullptr);
            });
            LAUNCH(h);
            k_sum_partials<<<1, 1024, 0, s.stream>>>(s.partials, s.grid_rows, s.scalars + 4);
            LAUNCH(h);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) { free_all(); return fail(TSC_ERR_CUDA, std::string("calculate_lnl: ") + cudaGetErrorString(e)); }
    }
    rc = allreduce(h, [](Shard& s) -> void* { return (void*)(s.scalars + 4); }, 1, ncclFloat64, ncclSum);
    if (!rc) rc = sync_all(h);
    free_all();
    if (rc) return rc;
    Shard& s = h->shards[0];
    CU(cudaSetDevice(s.dev));
    CU(cudaMemcpy(lnl, s.scalars + 4, sizeof(double), cudaMemcpyDeviceToHost));
    return TSC_OK;
}

// ------------------------------------------------------------------------------------------------- reassign
static int reassign_impl(tsc_handle* h, int method, double thresh, int initial, const int32_t* picks,
                         int32_t* nbest_rows, double* colsum, double* data) {
    if (method < 0 || method > 5) return fail(TSC_ERR_ARG, "Argument \"method\" should be one of (exclude, choose, average, conf, unique, all)");
    if (!initial && !h->em_done) return fail(TSC_ERR_STATE, "reassign(initial=False) needs em() first (self.z is None, model.py:659)");
    const int K = h->K;
    const bool compact = !h->rowmap.empty();
    std::vector<int32_t> picks_c, nbest_c;
    if (compact && picks) { picks_c.resize(h->n_rows); for (long long r = 0; r < h->n_rows; ++r) picks_c[r] = picks[h->rowmap[r]]; }
    if (compact && nbest_rows) nbest_c.resize(h->n_rows);
    const int32_t* picks_h = compact ? (picks ? picks_c.data() : nullptr) : picks;
    int32_t* nbest_h = compact ? (nbest_rows ? nbest_c.data() : nullptr) : nbest_rows;
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        int *picks_d = nullptr, *nbest_d = nullptr;
        double* data_d = nullptr;
        cudaError_t e = cudaSuccess;
        const size_t rb = sizeof(int) * std::max<long long>(s.n_rows, 1);
        if (picks_h && method == TSC_CHOOSE) {
            e = dev_malloc(&picks_d, rb);
            if (e == cudaSuccess) e = cudaMemcpyAsync(picks_d, picks_h + s.row_begin, sizeof(int) * s.n_rows, cudaMemcpyHostToDevice, s.stream);
            h->h2d += sizeof(int) * s.n_rows;
        }
        if (e == cudaSuccess && nbest_h) e = dev_malloc(&nbest_d, rb);
        if (e == cudaSuccess && data) e = dev_malloc(&data_d, sizeof(double) * std::max<long long>(s.nnz, 1));
        if (e == cudaSuccess && colsum) e = cudaMemsetAsync(s.colsum, 0, sizeof(double) * K, s.stream);
        if (e == cudaSuccess) {
            const double* ta = initial ? s.ones : s.pt_prev;
            const double* tu = initial ? s.ones : s.pi_prev;
            if (!data_d && !picks_d) {          // column sums / best-hit counts only: the stream serves them
                if (launch_reassign_sums(h, s, method, thresh, ta, tu, colsum ? s.colsum : nullptr, nbest_d)) e = cudaErrorUnknown;
            } else {
                ReassignArgs g{method, thresh, picks_d, nbest_d, colsum ? s.colsum : nullptr, data_d, nullptr};
                launch_rows(h->G, [&](auto gg) {
                    k_reassign_rows<decltype(gg)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), ta, tu, g);
                });
                LAUNCH(h);
                e = cudaGetLastError();
            }
        }
        if (e == cudaSuccess && nbest_h) { e = cudaMemcpyAsync(nbest_h + s.row_begin, nbest_d, sizeof(int) * s.n_rows, cudaMemcpyDeviceToHost, s.stream); h->d2h += sizeof(int) * s.n_rows; }
        if (e == cudaSuccess && data) { e = cudaMemcpyAsync(data + s.nnz_begin, data_d, sizeof(double) * s.nnz, cudaMemcpyDeviceToHost, s.stream); h->d2h += sizeof(double) * s.nnz; }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        if (picks_d) dev_free(picks_d);
        if (nbest_d) dev_free(nbest_d);
        if (data_d) dev_free(data_d);
        if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("reassign: ") + cudaGetErrorString(e));
    }
    if (compact && nbest_rows) {
        std::fill(nbest_rows, nbest_rows + h->n_rows_user, 0);
        for (long long r = 0; r < h->n_rows; ++r) nbest_rows[h->rowmap[r]] = nbest_c[r];
    }
    if (colsum) {
        ALLREDUCE(h, s.colsum, (size_t)K, ncclFloat64, ncclSum);
        int rc = sync_all(h);
        if (rc) return rc;
        Shard& s = h->shards[0];
        CU(cudaSetDevice(s.dev));
        return get_kvec(h, s, s.colsum, colsum);
    }
    return TSC_OK;
}

// per-read int arrays between the caller's read numbering and the compacted one (reads without entries are dropped)
static void rows_to_user(const tsc_handle* h, const std::vector<int32_t>& compact, int32_t* user) {
    std::fill(user, user + h->n_rows_user, 0);
    for (long long r = 0; r < h->n_rows; ++r) user[h->rowmap[r]] = compact[r];
}

extern "C" int tsc_report(tsc_handle* h, double thresh, int32_t final_method, int32_t* nbest_init, int32_t* nbest_final,
                          double* out6k) {
    if (!h || !out6k) return fail(TSC_ERR_ARG, "NULL argument");
    if (final_method < 0 || final_method > 5) return fail(TSC_ERR_ARG, "Argument \"method\" should be one of (exclude, choose, average, conf, unique, all)");
    if (!h->em_done) return fail(TSC_ERR_STATE, "the report needs em() first");
    const int K = h->K;
    const bool compact = !h->rowmap.empty();
    std::vector<int32_t> nbi_c, nbf_c;
    if (compact && nbest_init) nbi_c.resize(h->n_rows);
    if (compact && nbest_final) nbf_c.resize(h->n_rows);
    int32_t* nbi_h = compact ? (nbest_init ? nbi_c.data() : nullptr) : nbest_init;
    int32_t* nbf_h = compact ? (nbest_final ? nbf_c.data() : nullptr) : nbest_final;
    std::vector<double*> out_d(h->shards.size(), nullptr);
    auto free_all = [&]() { for (size_t i = 0; i < out_d.size(); ++i) if (out_d[i]) { cudaSetDevice(h->shards[i].dev); dev_free(out_d[i]); } };
    for (size_t i = 0; i < h->shards.size(); ++i) {
        Shard& s = h->shards[i];
        CU(cudaSetDevice(s.dev));
        int *nbi_d = nullptr, *nbf_d = nullptr;
        const size_t rb = sizeof(int) * std::max<long long>(s.n_rows, 1);
        cudaError_t e = dev_malloc(&out_d[i], sizeof(double) * 6 * K);
        if (e == cudaSuccess) e = cudaMemsetAsync(out_d[i], 0, sizeof(double) * 6 * K, s.stream);
        if (e == cudaSuccess && nbi_h) e = dev_malloc(&nbi_d, rb);
        if (e == cudaSuccess && nbf_h) e = dev_malloc(&nbf_d, rb);
        if (e == cudaSuccess) {
            ReportArgs g{thresh, final_method, nbi_d, nbf_d, out_d[i], K};
            launch_rows(h->G, [&](auto gg) {
                k_report_rows<decltype(gg)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), s.pt_prev, s.pi_prev, g);
            });
            LAUNCH(h);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess && nbi_h) { e = cudaMemcpyAsync(nbi_h + s.row_begin, nbi_d, sizeof(int) * s.n_rows, cudaMemcpyDeviceToHost, s.stream); h->d2h += sizeof(int) * s.n_rows; }
        if (e == cudaSuccess && nbf_h) { e = cudaMemcpyAsync(nbf_h + s.row_begin, nbf_d, sizeof(int) * s.n_rows, cudaMemcpyDeviceToHost, s.stream); h->d2h += sizeof(int) * s.n_rows; }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        if (nbi_d) dev_free(nbi_d);
        if (nbf_d) dev_free(nbf_d);
        if (e != cudaSuccess) { free_all(); return fail(TSC_ERR_CUDA, std::string("report: ") + cudaGetErrorString(e)); }
    }
    for (size_t i = 0; i < h->shards.size(); ++i) h->shards[i].xchg_ptr = out_d[i];
    {
        const int rc_ar = allreduce(h, [](Shard& s) -> void* { return s.xchg_ptr; }, (size_t)6 * K, ncclFloat64, ncclSum);
        if (rc_ar) { free_all(); return rc_ar; }
    }
    int rc = sync_all(h);
    Shard& s0 = h->shards[0];
    for (int v = 0; v < 6 && !rc; ++v) {
        cudaSetDevice(s0.dev);
        rc = get_kvec(h, s0, out_d[0] + (size_t)v * K, out6k + (size_t)v * K);
    }
    free_all();
    if (rc) return rc;
    for (int j = 0; j < K; ++j) out6k[(size_t)K + j] = (double)h->pos_count[j];      // init_aligned: known since construction
    if (compact && nbest_init) rows_to_user(h, nbi_c, nbest_init);
    if (compact && nbest_final) rows_to_user(h, nbf_c, nbest_final);
    return TSC_OK;
}

extern "C" int tsc_choose_ties_colsum(tsc_handle* h, int32_t initial, const int32_t* nbest, const int32_t* picks, double* colsum) {
    if (!h || !nbest || !picks || !colsum) return fail(TSC_ERR_ARG, "NULL argument");
    if (!initial && !h->em_done) return fail(TSC_ERR_STATE, "reassign(initial=False) needs em() first");
    const int K = h->K;
    const bool compact = !h->rowmap.empty();
    std::vector<int32_t> nb_c, pk_c;
    if (compact) {
        nb_c.resize(h->n_rows); pk_c.resize(h->n_rows);
        for (long long r = 0; r < h->n_rows; ++r) { nb_c[r] = nbest[h->rowmap[r]]; pk_c[r] = picks[h->rowmap[r]]; }
    }
    const int32_t* nb_h = compact ? nb_c.data() : nbest;
    const int32_t* pk_h = compact ? pk_c.data() : picks;
    for (auto& s : h->shards) {
        CU(cudaSetDevice(s.dev));
        int *nb_d = nullptr, *pk_d = nullptr;
        const size_t rb = sizeof(int) * std::max<long long>(s.n_rows, 1);
        cudaError_t e = dev_malloc(&nb_d, rb);
        if (e == cudaSuccess) e = dev_malloc(&pk_d, rb);
        if (e == cudaSuccess) e = cudaMemcpyAsync(nb_d, nb_h + s.row_begin, sizeof(int) * s.n_rows, cudaMemcpyHostToDevice, s.stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(pk_d, pk_h + s.row_begin, sizeof(int) * s.n_rows, cudaMemcpyHostToDevice, s.stream);
        h->h2d += 2 * sizeof(int) * s.n_rows;
        if (e == cudaSuccess) e = cudaMemsetAsync(s.colsum, 0, sizeof(double) * K, s.stream);
        if (e == cudaSuccess) {
            const double* ta = initial ? s.ones : s.pt_prev;
            const double* tu = initial ? s.ones : s.pi_prev;
            launch_rows(h->G, [&](auto gg) {
                k_choose_ties_rows<decltype(gg)::value><<<s.grid_rows, 512, 0, s.stream>>>(csr_of(s), ta, tu, nb_d, pk_d, s.colsum);
            });
            LAUNCH(h);
            e = cudaGetLastError();
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(s.stream);
        if (nb_d) dev_free(nb_d);
        if (pk_d) dev_free(pk_d);
        if (e != cudaSuccess) return fail(TSC_ERR_CUDA, std::string("choose ties: ") + cudaGetErrorString(e));
    }
    ALLREDUCE(h, s.colsum, (size_t)K, ncclFloat64, ncclSum);
    int rc = sync_all(h);
    if (rc) return rc;
    Shard& s0 = h->shards[0];
    CU(cudaSetDevice(s0.dev));
    return get_kvec(h, s0, s0.colsum, colsum);
}

// ------------------------------------------------------------------------------------------------- host RNG (choose)
// reassign('choose') breaks ties with np.random.choice(range(a, b)) per read, in read order (sparse_plus.py:146-153), i.e.
// one bounded draw of numpy's legacy global generator each: MT19937 words, masked with the smallest 2^k - 1 >= n - 1 and
// rejected while above n - 1.  The draws are the reference's contract (same seed -> same assignments), and 25 M of them
// through numpy's array path were two thirds of the one-pass report's wall time; this walks the same generator state in C.
// key624 / pos are RandomState.get_state()'s; both are updated for set_state().  Host code only, no device involved.
template <typename CountT>
static int mt19937_draw(uint32_t* key624, int32_t* pos, const CountT* counts, int64_t n, int32_t* picks) {
    if (!key624 || !pos || (n > 0 && (!counts || !picks))) return fail(TSC_ERR_ARG, "NULL argument");
    if (*pos < 0 || *pos > 624) return fail(TSC_ERR_ARG, "MT19937 position out of range");
    constexpr int N = 624, M = 397;
    constexpr uint32_t kMatrixA = 0x9908b0dfu, kUpper = 0x80000000u, kLower = 0x7fffffffu;
    int p = *pos;
    auto regenerate = [&]() {
        int k = 0;
        uint32_t y;
        for (; k < N - M; ++k) {
            y = (key624[k] & kUpper) | (key624[k + 1] & kLower);
            key624[k] = key624[k + M] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
        }
        for (; k < N - 1; ++k) {
            y = (key624[k] & kUpper) | (key624[k + 1] & kLower);
            key624[k] = key624[k + (M - N)] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
        }
        y = (key624[N - 1] & kUpper) | (key624[0] & kLower);
        key624[N - 1] = key624[M - 1] ^ (y >> 1) ^ ((0u - (y & 1u)) & kMatrixA);
        p = 0;
    };
    auto next32 = [&]() -> uint32_t {
        if (p == N) regenerate();
        uint32_t y = key624[p++];
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    };
    for (int64_t i = 0; i < n; ++i) {
        const long long c = (long long)counts[i];
        if (c <= 1) { picks[i] = 0; continue; }                 // nothing to choose: consumes nothing (like numpy for n = 1)
        if (c > 0x7fffffffLL) { *pos = p; return fail(TSC_ERR_ARG, "counts must be below 2^31"); }
        const uint32_t rng = (uint32_t)(c - 1);
        uint32_t mask = rng;
        mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
        uint32_t v;
        do { v = next32() & mask; } while (v > rng);
        picks[i] = (int32_t)v;
    }
    *pos = p;
    return TSC_OK;
}

extern "C" int tsc_mt19937_draw_picks(uint32_t* key624, int32_t* pos, const int64_t* counts, int64_t n, int32_t* picks) {
    return mt19937_draw<int64_t>(key624, pos, counts, n, picks);
}
extern "C" int tsc_mt19937_draw_rows(uint32_t* key624, int32_t* pos, const int32_t* nbest, int64_t n, int32_t* picks) {
    return mt19937_draw<int32_t>(key624, pos, nbest, n, picks);
}

extern "C" int tsc_reassign_nbest(tsc_handle* h, int32_t initial, int32_t* nbest_rows) {
    if (!h || !nbest_rows) return fail(TSC_ERR_ARG, "NULL argument");
    return reassign_impl(h, TSC_EXCLUDE, 0.0, initial, nullptr, nbest_rows, nullptr, nullptr);
}
extern "C" int tsc_reassign_colsum(tsc_handle* h, int32_t method, double thresh, int32_t initial, const int32_t* picks, double* colsum) {
    if (!h || !colsum) return fail(TSC_ERR_ARG, "NULL argument");
    return reassign_impl(h, method, thresh, initial, picks, nullptr, colsum, nullptr);
}
extern "C" int tsc_reassign_data(tsc_handle* h, int32_t method, double thresh, int32_t initial, const int32_t* picks, double* data) {
    if (!h || !data) return fail(TSC_ERR_ARG, "NULL argument");
    return reassign_impl(h, method, thresh, initial, picks, nullptr, nullptr, data);
}
