# -*- coding: utf-8 -*-
"""`TelescopeLikelihood` backed by libtelescope_b200.so -- same constructor, methods and attributes as the reference
class (telescope/utils/model.py:631-865), with the EM loop, the posterior, the log-likelihood and the six
reassignment modes running as CUDA kernels on B200.  There is no CPU path: constructing the class without the built
library or without a GPU raises.

What stays on the host, by design:
  * the 65536-entry-at-most table Q(s) = expm1((s * (1/max_score)) * 100.) is evaluated with numpy exactly as
    model.py:652-653 does, so the device Q is bit-identical to the reference's;
  * `reassign('choose')` draws its tie-breaks from the global numpy RNG (the device reports how many best hits each
    read has), so `np.random.seed(...)` in telescope_assign.run / telescope_resume.run keeps its meaning.
"""
import ctypes as C
import logging as lg

import numpy as np

from . import _abi
from .sparse_plus import csr_matrix_plus as csr_matrix
from .sparse_plus import draw_picks, draw_row_picks          # noqa: F401

_INT_DTYPE = {"exclude": np.int8, "choose": np.int8, "unique": np.uint8, "all": np.uint8}


def _nccl_env_tuning():
    """Opt-in (TELESCOPE_B200_NCCL_TUNE=1) NCCL settings for the NCCL transport: the only collective on this path is a
    latency-bound all-reduce of K doubles, so communicator SETUP cost is what matters -- without NVLS (multicast)
    setup and with two channels ncclCommInitRank takes about half as long on an 8 x B200 box and the per-iteration
    time is unchanged (profiles/r1_nccl_init.md).  The variables are process-global (every later communicator of the
    process inherits them), which is why the library never sets them on its own."""
    import os
    if os.environ.get("TELESCOPE_B200_NCCL_TUNE", "0") != "1":
        return
    os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
    os.environ.setdefault("NCCL_MAX_NCHANNELS", "2")


def trim_memory():
    """Hand the device memory kept from destroyed models back to the driver (the library keeps a bounded cache of blocks
    between the models of one process: TELESCOPE_B200_CACHE_GB, default 96, 0 = off)."""
    lib = _abi.load()
    lib.tsc_trim_memory.restype = None
    lib.tsc_trim_memory()


class DistInfo(object):
    """How this process takes part in a multi-process run (one process per GPU, e.g. under torchrun).

    transport "peer" (default on one node): the ranks' K-length sums travel through CUDA-IPC mapped peer memory inside
    the update kernel; `allgather(tag, bytes) -> [bytes]` swaps the IPC handles at construction.  transport "nccl":
    `nccl_id` is rank 0's 128-byte id."""

    def __init__(self, n_procs=1, proc_rank=0, nccl_id=None, allgather=None, transport=None):
        self.n_procs, self.proc_rank, self.nccl_id, self.allgather = n_procs, proc_rank, nccl_id, allgather
        self.transport = transport or ("nccl" if nccl_id is not None else "peer")


class TelescopeLikelihood(object):

    def __init__(self, score_matrix, opts, devices=None, dist=None, max_score=None, kernel="auto", replicas=0,
                 smem_table_cols=-1, permute_columns=False, transport="auto"):
        """score_matrix: csr_matrix (uint16) N reads x K loci; opts: em_epsilon, max_iter, pi_prior, theta_prior.

        devices: list of CUDA ordinals driven from this process (default [0]).  dist: DistInfo when this process
        holds only its block of reads; `max_score` must then be the global maximum score.  transport: "auto", "peer"
        (exchange through peer-mapped GPU memory inside the update kernel; one node) or "nccl".
        """
        self._lib = _abi.load()
        self._h = None
        self.raw_scores = score_matrix
        if score_matrix.nnz >= 2 ** 31 and score_matrix.indptr.dtype != np.int64:
            raise ValueError("indptr must be int64 for >= 2^31 entries")
        if dist is not None and dist.n_procs > 1 and max_score is None:
            # every rank must build Q = expm1(100*s/max) with the SAME maximum (model.py:640 takes it over all reads)
            raise ValueError("max_score (the global maximum score) is required when the reads are split over processes")
        self.max_score = score_matrix.max() if max_score is None else max_score        # model.py:640
        self.N, self.K = score_matrix.shape                                               # model.py:643
        self.scale_factor = 100.                                                          # model.py:652
        ms = int(self.max_score)
        if ms <= 0:
            raise ValueError("score matrix has no positive scores")
        # Q lookup table, evaluated as the reference evaluates Q (model.py:653, sparse_plus.py:89-91)
        self._lut = np.expm1((np.arange(ms + 1, dtype=np.float64) * (1. / ms)) * self.scale_factor)

        self.epsilon = opts.em_epsilon                                                    # model.py:661-662
        self.max_iter = opts.max_iter
        self.pi = np.repeat(1. / self.K, self.K)                                          # model.py:667
        self.pi_init = None
        self.theta = np.repeat(1. / self.K, self.K)                                       # model.py:673
        self.theta_init = None
        self.lnl = float('inf')                                                           # model.py:683
        self.pi_prior = opts.pi_prior                                                     # model.py:686-687
        self.theta_prior = opts.theta_prior
        self._z = None
        self._z_stale = True
        self._Q = None
        self._Y = None
        self._w = None
        self.diffs = []
        self.lnls = []
        self.n_iter = 0
        self.converged = False

        indptr = np.ascontiguousarray(score_matrix.indptr)
        if indptr.dtype not in (np.int32, np.int64):
            indptr = indptr.astype(np.int64)
        indices = np.ascontiguousarray(score_matrix.indices, dtype=np.int32)
        raw = np.ascontiguousarray(score_matrix.data, dtype=np.uint16)
        if score_matrix.data.dtype != np.uint16 and raw.size and not np.array_equal(raw, score_matrix.data):
            raise ValueError("scores must be integers in [0, 65535] (the reference stores uint16, model.py:300)")
        self._indptr, self._indices = indptr, indices

        cfg = _abi.TscConfig()
        self._lib.tsc_config_default(C.byref(cfg))
        devices = [0] if devices is None else list(devices)
        self._dev_arr = (C.c_int32 * len(devices))(*devices)
        cfg.n_local_devices = len(devices)
        cfg.device_ids = C.cast(self._dev_arr, C.POINTER(C.c_int32))
        self._nccl_id = self._peer_handles = None
        cfg.transport = _abi.TRANSPORTS[transport]
        if dist is not None and dist.n_procs > 1:
            cfg.n_procs, cfg.proc_rank = dist.n_procs, dist.proc_rank
            if transport == "auto":
                cfg.transport = _abi.TRANSPORTS[dist.transport]
            if cfg.transport == _abi.TRANSPORTS["peer"]:
                # every rank allocates its exchange buffer, the ranks swap the CUDA IPC handles (dist.allgather)
                if len(devices) != 1 or dist.allgather is None:
                    raise ValueError("the peer transport between processes takes one device per process and DistInfo.allgather")
                buf, handle = C.c_void_p(), C.create_string_buffer(64)
                _abi.check(self._lib.tsc_peer_buffer_create(devices[0], self.K, dist.n_procs, C.byref(buf), handle))
                try:
                    blobs = dist.allgather("ipc", handle.raw)
                except Exception:
                    self._lib.tsc_peer_buffer_free(buf)
                    raise
                self._peer_handles = C.create_string_buffer(b"".join(blobs), 64 * dist.n_procs)
                cfg.peer_buffer, cfg.peer_handles = buf, C.cast(self._peer_handles, C.c_void_p)
            else:
                if dist.nccl_id is None:
                    raise ValueError("the NCCL transport needs DistInfo.nccl_id (dist.rendezvous(transport='nccl'))")
                self._nccl_id = C.create_string_buffer(dist.nccl_id, 128)
                cfg.nccl_id = C.cast(self._nccl_id, C.c_void_p)
        if cfg.transport != _abi.TRANSPORTS["peer"] and (len(devices) > 1 or (dist is not None and dist.n_procs > 1)):
            _nccl_env_tuning()
            path = _abi.find_nccl()
            if path:
                self._lib.tsc_set_nccl_path(path.encode())
        cfg.kernel = _abi.KERNELS[kernel]
        cfg.replicas = replicas
        cfg.smem_table_cols = smem_table_cols
        cfg.permute_columns = 1 if permute_columns else 0
        h = C.c_void_p()
        import time as _time
        _t0 = _time.perf_counter()
        _abi.check(self._lib.tsc_create(
            C.byref(h), C.byref(cfg), self.N, self.K, int(indices.size),
            indptr.ctypes.data_as(C.c_void_p), indptr.dtype.itemsize, _abi._p(indices, C.c_int32),
            _abi._p(raw, C.c_uint16), _abi._p(self._lut, C.c_double), int(self._lut.size),
            float(self.pi_prior), float(self.theta_prior)))
        self._h = h
        self.create_seconds = _time.perf_counter() - _t0          # diagnostics: time inside tsc_create
        self.create_laps = self._lib.tsc_create_laps(self._h).decode()

        sc = np.zeros(5)
        pisum0 = np.zeros(self.K)
        _abi.check(self._lib.tsc_get_constants(self._h, _abi._p(sc, C.c_double), _abi._p(pisum0, C.c_double)))
        self._total_wt, self._ambig_wt, self._max_wt = sc[0], sc[1], sc[2]                # model.py:691-692
        self._unique_wt = self._total_wt - self._ambig_wt                                 # model.py:693
        self._pi_prior_wt, self._theta_prior_wt = sc[3], sc[4]                            # model.py:696-697
        self._pisum0 = np.asmatrix(pisum0)                                                # model.py:699
        lg.debug('done initializing model')

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self._lib.tsc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ lazily materialised attributes
    def _wrap(self, data, dtype=None):
        if dtype is not None:
            data = data.astype(dtype)
        return csr_matrix((data, self._indices.copy(), self._indptr.copy()), shape=(self.N, self.K))

    @property
    def Q(self):
        if self._Q is None:
            q = np.empty(self._indices.size)
            _abi.check(self._lib.tsc_get_q(self._h, _abi._p(q, C.c_double)))
            self._Q = self._wrap(q)
        return self._Q

    def _row_info(self):
        if self._Y is None:
            y = np.zeros(self.N, dtype=np.uint8)
            w = np.zeros(self.N)
            _abi.check(self._lib.tsc_get_row_info(self._h, _abi._p(y, C.c_uint8), _abi._p(w, C.c_double)))
            self._Y, self._w = y.reshape(-1, 1), w
        return self._Y, self._w

    @property
    def Y(self):
        return self._row_info()[0]

    @property
    def _weights(self):
        import scipy.sparse
        return scipy.sparse.coo_matrix(self._row_info()[1].reshape(-1, 1))

    @property
    def _yslice(self):
        return self.Y[:, 0].nonzero()[0]

    @property
    def z(self):
        """Posterior of the last E-step (model.py:795).  Pulled from the device on first access."""
        if self.n_iter == 0:
            return None
        if self._z_stale:
            d = np.empty(self._indices.size)
            _abi.check(self._lib.tsc_get_z(self._h, 0, _abi._p(d, C.c_double)))
            self._z = self._finish_z(d)
            self._z_stale = False
        return self._z

    def _finish_z(self, data):
        z = self._wrap(data)
        z.eliminate_zeros()          # the reference's sparse add drops exact zeros (model.py:720)
        return z

    def _entry_data(self, m):
        """Data of a matrix whose structure is a subset of raw_scores', expanded to raw_scores' entry order."""
        m = csr_matrix(m)
        if m.shape != (self.N, self.K):
            raise ValueError("matrix shape %s does not match (%d, %d)" % (m.shape, self.N, self.K))
        if m.nnz == self._indices.size and np.array_equal(m.indptr, self._indptr) and np.array_equal(m.indices, self._indices):
            return np.ascontiguousarray(m.data, dtype=np.float64)
        m.sort_indices()
        K = np.int64(self.K)
        full = np.repeat(np.arange(self.N, dtype=np.int64), np.diff(self._indptr)) * K + self._indices
        sub = np.repeat(np.arange(self.N, dtype=np.int64), np.diff(m.indptr)) * K + m.indices
        pos = np.searchsorted(full, sub)
        if pos.size and (pos.max() >= full.size or not np.array_equal(full[pos], sub)):
            raise ValueError("matrix has entries outside the score matrix's structure")
        out = np.zeros(full.size)
        out[pos] = m.data
        return out

    # ------------------------------------------------------------------ model.py:702-760
    def estep(self, pi, theta):
        lg.debug('started e-step')
        pi, theta = _abi.as_f64(pi), _abi.as_f64(theta)
        d = np.empty(self._indices.size)
        _abi.check(self._lib.tsc_estep(self._h, _abi._p(pi, C.c_double), _abi._p(theta, C.c_double), _abi._p(d, C.c_double)))
        return self._finish_z(d)

    def mstep(self, z):
        lg.debug('started m-step')
        zd = self._entry_data(z)
        pi_hat, theta_hat = np.empty(self.K), np.empty(self.K)
        _abi.check(self._lib.tsc_mstep(self._h, _abi._p(zd, C.c_double), _abi._p(pi_hat, C.c_double), _abi._p(theta_hat, C.c_double)))
        return pi_hat, theta_hat

    def calculate_lnl(self, z, pi, theta):
        lg.debug('started lnl')
        zd = self._entry_data(z)
        pi, theta = _abi.as_f64(pi), _abi.as_f64(theta)
        out = C.c_double(0)
        _abi.check(self._lib.tsc_calculate_lnl(self._h, _abi._p(zd, C.c_double), _abi._p(pi, C.c_double),
                                                _abi._p(theta, C.c_double), C.byref(out)))
        lg.debug('completed lnl')
        return out.value

    # ------------------------------------------------------------------ model.py:762-806
    def em(self, use_likelihood=False, loglev=lg.WARNING, save_memory=True):
        msgD = 'Iteration {:d}, diff={:.5g}'
        msgL = 'Iteration {:d}, lnl= {:.5e}, diff={:.5g}'
        T = max(1, int(self.max_iter))
        # the device loop continues from the current self.pi / self.theta (model.py:773)
        pi, theta = _abi.as_f64(self.pi), _abi.as_f64(self.theta)
        _abi.check(self._lib.tsc_set_params(self._h, _abi._p(pi, C.c_double), _abi._p(theta, C.c_double)))
        diffs, lnls = np.zeros(T), np.zeros(T)
        n_iter, conv, lnl = C.c_int32(0), C.c_int32(0), C.c_double(0)
        _abi.check(self._lib.tsc_em(self._h, T, float(self.epsilon), 1 if use_likelihood else 0,
                                     _abi._p(diffs, C.c_double), _abi._p(lnls, C.c_double),
                                     C.byref(n_iter), C.byref(conv), C.byref(lnl)))
        inum, converged = n_iter.value, bool(conv.value)
        self.diffs, self.lnls = diffs[:inum].tolist(), (lnls[:inum].tolist() if use_likelihood else [])
        for i in range(inum):
            if use_likelihood:
                lg.log(loglev, msgL.format(i + 1, lnls[i], diffs[i]))
            else:
                lg.log(loglev, msgD.format(i + 1, diffs[i]))
        pi, theta, pi0, theta0 = (np.empty(self.K) for _ in range(4))
        _abi.check(self._lib.tsc_get_params(self._h, *(_abi._p(a, C.c_double) for a in (pi, theta, pi0, theta0))))
        self.pi, self.theta = pi, theta
        self.pi_init, self.theta_init = pi0, theta0                                     # model.py:776-778
        self.lnl = lnl.value
        self.n_iter, self.converged = inum, converged
        self._z_stale = True
        _con = 'converged' if converged else 'terminated'
        lg.log(loglev, 'EM {:s} after {:d} iterations.'.format(_con, inum))
        lg.log(loglev, 'Final log-likelihood: {:f}.'.format(self.lnl))
        return

    def kernel_times_ms(self):
        n = C.c_int32(0)
        buf = np.zeros(max(1, self.n_iter), dtype=np.float32)
        _abi.check(self._lib.tsc_get_kernel_times(self._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)]

    def tail_times_ms(self):
        """Per-iteration device time of the iteration's tail (exchange between GPUs + update + loop control)."""
        n = C.c_int32(0)
        buf = np.zeros(max(1, self.n_iter), dtype=np.float32)
        _abi.check(self._lib.tsc_get_tail_times(self._h, buf.ctypes.data_as(C.POINTER(C.c_float)), buf.size, C.byref(n)))
        return buf[:min(n.value, buf.size)]

    def em_device_ms(self):
        ms = C.c_float(0)
        _abi.check(self._lib.tsc_get_em_device_ms(self._h, C.byref(ms)))
        return ms.value

    def allreduce(self, values, op="sum"):
        """Sum / max of a few host doubles over all ranks of this model's communicator (identity on one GPU)."""
        buf = np.ascontiguousarray(np.atleast_1d(values), dtype=np.float64).copy()
        _abi.check(self._lib.tsc_allreduce_f64(self._h, _abi._p(buf, C.c_double), buf.size, {"sum": 0, "max": 1}[op]))
        return buf

    def time_pass(self, which, reps=5):
        """Mean device time (ms) of one pass on local shard 0: 'fused', 'estep', 'lnl' or 'reassign' (diagnostic)."""
        ms = C.c_float(0)
        _abi.check(self._lib.tsc_time_pass(self._h, {"fused": 0, "estep": 1, "lnl": 2, "reassign": 3}[which], reps, C.byref(ms)))
        return ms.value

    def transport(self):
        """'peer' or 'nccl': how this model's GPUs exchange the per-locus sums."""
        t = C.c_int32(0)
        _abi.check(self._lib.tsc_get_transport(self._h, C.byref(t)))
        return {v: k for k, v in _abi.TRANSPORTS.items()}[t.value]

    def layout_stats(self):
        """Device layout of local shard 0 (diagnostic): the clustered slice stream and the residual CSR."""
        buf = (C.c_int64 * 8)()
        _abi.check(self._lib.tsc_get_layout_stats(self._h, buf))
        names = ("stream_bytes", "slices", "stream_reads", "stream_entries", "residual_reads", "residual_entries",
                 "stream_ctas", "long_reads")
        return dict(zip(names, (int(v) for v in buf)))

    def counters(self):
        a, b, c = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        _abi.check(self._lib.tsc_get_counters(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return {"launches": a.value, "h2d_bytes": b.value, "d2h_bytes": c.value}

    # ------------------------------------------------------------------ model.py:808-865
    def _check_method(self, method):
        if method not in ['exclude', 'choose', 'average', 'conf', 'unique', 'all']:
            raise ValueError('Argument "method" should be one of (exclude, choose, average, conf, unique, all)')
        return _abi.METHODS[method]

    def _picks(self, method, initial):
        """Tie-breaks for 'choose': one draw per read with more than one best hit, in read order."""
        if method != 'choose':
            return None
        nbest = np.zeros(self.N, dtype=np.int32)
        _abi.check(self._lib.tsc_reassign_nbest(self._h, 1 if initial else 0, _abi._p(nbest, C.c_int32)))
        return draw_row_picks(nbest)

    def reassign(self, method, thresh=0.9, initial=False):
        """Assignment matrix, as the reference returns it (int8 / uint8 / float64 csr_matrix)."""
        m = self._check_method(method)
        picks = self._picks(method, initial)
        d = np.empty(self._indices.size)
        _abi.check(self._lib.tsc_reassign_data(self._h, m, float(thresh), 1 if initial else 0,
                                                _abi._p(picks, C.c_int32) if picks is not None else None,
                                                _abi._p(d, C.c_double)))
        out = self._wrap(d, _INT_DTYPE.get(method))
        out.eliminate_zeros()
        return out

    def report_colsums(self, conf_prob=0.9, final_method='exclude'):
        """Every column sum `Telescope.output_report` needs (model.py:432-458), from ONE pass over the matrix plus a
        small pass per 'choose' column.  Consumes the global numpy RNG in the reference's order: the draws of
        init_best_random first, then those of the counts column when final_method is 'choose'."""
        m = self._check_method(final_method)
        K, N = self.K, self.N
        out = np.empty(6 * K)
        nb_i = np.zeros(N, dtype=np.int32)
        nb_f = np.zeros(N, dtype=np.int32) if final_method == 'choose' else None
        _abi.check(self._lib.tsc_report(self._h, float(conf_prob), m, _abi._p(nb_i, C.c_int32),
                                         _abi._p(nb_f, C.c_int32) if nb_f is not None else None, _abi._p(out, C.c_double)))
        out = out.reshape(6, K)

        def ties(nbest, initial):
            picks = draw_row_picks(nbest)
            extra = np.zeros(K)
            if nbest.size and int(nbest.max()) > 1:
                _abi.check(self._lib.tsc_choose_ties_colsum(self._h, 1 if initial else 0, _abi._p(nbest, C.c_int32),
                                                            _abi._p(picks, C.c_int32), _abi._p(extra, C.c_double)))
            return extra

        res = {
            'final_conf': out[0].copy(),
            'init_aligned': np.rint(out[1]).astype(np.uint64),
            'unique_count': np.rint(out[2]).astype(np.uint64),
            'init_best': np.rint(out[3]).astype(np.int64),
            'init_best_random': np.rint(out[3] + ties(nb_i, True)).astype(np.int64),
            'init_best_avg': out[4].copy(),
        }
        final = out[5]
        if final_method == 'choose':
            final = final + ties(nb_f, False)
        if final_method in ('exclude', 'choose'):
            final = np.rint(final).astype(np.int64)
        elif final_method in ('unique', 'all'):
            final = np.rint(final).astype(np.uint64)
        res['final'] = final
        return res

    def reassign_colsum(self, method, thresh=0.9, initial=False):
        """`reassign(method, thresh, initial).sum(0).A1` without moving the N x K matrix off the device."""
        m = self._check_method(method)
        picks = self._picks(method, initial)
        out = np.empty(self.K)
        _abi.check(self._lib.tsc_reassign_colsum(self._h, m, float(thresh), 1 if initial else 0,
                                                  _abi._p(picks, C.c_int32) if picks is not None else None,
                                                  _abi._p(out, C.c_double)))
        if method in ('exclude', 'choose'):
            return np.rint(out).astype(np.int64)      # scipy sums int8 into int64
        if method in ('unique', 'all'):
            return np.rint(out).astype(np.uint64)     # and uint8 into uint64
        return out
