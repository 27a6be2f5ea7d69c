# -*- coding: utf-8 -*-
"""ctypes binding of libtelescope_b200.so (include/telescope_b200.h).  No torch, no fallback.

This is the stub a maintainer of the reference would add next to telescope/utils/model.py to call the CUDA
library; INTEGRATION.md walks through it.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TELESCOPE_B200_LIB") or os.path.join(_HERE, "libtelescope_b200.so")

TSC_OK, TSC_ERR_ARG, TSC_ERR_CUDA, TSC_ERR_NCCL, TSC_ERR_STATE, TSC_ERR_ALLOC = range(6)
METHODS = {"exclude": 0, "choose": 1, "average": 2, "conf": 3, "unique": 4, "all": 5}
KERNELS = {"auto": 0, "rows": 1, "tiles": 2, "ell": 3}
TRANSPORTS = {"auto": 0, "peer": 1, "nccl": 2}

# every symbol the header declares; tests check the library exports exactly these
SYMBOLS = (
    "tsc_abi_version", "tsc_last_error", "tsc_device_count", "tsc_set_nccl_path", "tsc_nccl_unique_id",
    "tsc_config_default", "tsc_create", "tsc_destroy", "tsc_get_constants", "tsc_get_row_info", "tsc_get_q",
    "tsc_em", "tsc_get_kernel_times", "tsc_get_counters", "tsc_get_params", "tsc_set_params", "tsc_estep",
    "tsc_mstep", "tsc_calculate_lnl", "tsc_get_z", "tsc_reassign_nbest", "tsc_reassign_colsum",
    "tsc_reassign_data", "tsc_get_em_device_ms", "tsc_pinned_alloc", "tsc_pinned_free", "tsc_time_pass", "tsc_allreduce_f64", "tsc_create_laps", "tsc_report", "tsc_choose_ties_colsum", "tsc_get_layout_stats", "tsc_peer_buffer_create", "tsc_peer_buffer_free", "tsc_get_transport", "tsc_get_tail_times", "tsc_trim_memory", "tsc_mt19937_draw_picks", "tsc_mt19937_draw_rows",
)


class TscConfig(C.Structure):
    _fields_ = [
        ("n_local_devices", C.c_int32),
        ("device_ids", C.POINTER(C.c_int32)),
        ("n_procs", C.c_int32),
        ("proc_rank", C.c_int32),
        ("nccl_id", C.c_void_p),
        ("kernel", C.c_int32),
        ("replicas", C.c_int32),
        ("smem_table_cols", C.c_int32),
        ("smem_acc_cols", C.c_int32),
        ("permute_columns", C.c_int32),
        ("transport", C.c_int32),
        ("peer_buffer", C.c_void_p),
        ("peer_handles", C.c_void_p),
        ("reserved", C.c_int32 * 4),
    ]


class TelescopeCudaError(RuntimeError):
    pass


_lib = None


def _p(arr, ctype):
    return arr.ctypes.data_as(C.POINTER(ctype))


def load():
    """dlopen the library and declare its prototypes.  Fails loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TelescopeCudaError(
            "%s is missing: build it with `python -m telescope_b200.build` (needs nvcc). "
            "telescope_b200 has no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp, ip = C.POINTER(C.c_double), C.POINTER(C.c_int32)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.tsc_abi_version.restype = C.c_int
    lib.tsc_last_error.restype = C.c_char_p
    lib.tsc_device_count.argtypes = [ip]
    lib.tsc_set_nccl_path.argtypes = [C.c_char_p]
    lib.tsc_nccl_unique_id.argtypes = [vp]
    lib.tsc_config_default.argtypes = [C.POINTER(TscConfig)]
    lib.tsc_config_default.restype = None
    lib.tsc_create.argtypes = [C.POINTER(vp), C.POINTER(TscConfig), i64, i32, i64, vp, i32, ip,
                               C.POINTER(C.c_uint16), dp, i32, dbl, dbl]
    lib.tsc_destroy.argtypes = [vp]
    lib.tsc_destroy.restype = None
    lib.tsc_get_constants.argtypes = [vp, dp, dp]
    lib.tsc_get_row_info.argtypes = [vp, C.POINTER(C.c_uint8), dp]
    lib.tsc_get_q.argtypes = [vp, dp]
    lib.tsc_em.argtypes = [vp, i32, dbl, i32, dp, dp, ip, ip, dp]
    lib.tsc_get_kernel_times.argtypes = [vp, C.POINTER(C.c_float), i32, ip]
    lib.tsc_get_tail_times.argtypes = [vp, C.POINTER(C.c_float), i32, ip]
    lib.tsc_get_em_device_ms.argtypes = [vp, C.POINTER(C.c_float)]
    lib.tsc_pinned_alloc.argtypes = [C.c_uint64]
    lib.tsc_pinned_alloc.restype = C.c_void_p
    lib.tsc_pinned_free.argtypes = [vp]
    lib.tsc_pinned_free.restype = None
    lib.tsc_report.argtypes = [vp, dbl, i32, ip, ip, dp]
    lib.tsc_choose_ties_colsum.argtypes = [vp, i32, ip, ip, dp]
    lib.tsc_create_laps.argtypes = [vp]
    lib.tsc_create_laps.restype = C.c_char_p
    lib.tsc_allreduce_f64.argtypes = [vp, dp, i32, i32]
    lib.tsc_time_pass.argtypes = [vp, i32, i32, C.POINTER(C.c_float)]
    lib.tsc_peer_buffer_create.argtypes = [i32, i32, i32, C.POINTER(vp), vp]
    lib.tsc_peer_buffer_free.argtypes = [vp]
    lib.tsc_peer_buffer_free.restype = None
    lib.tsc_get_transport.argtypes = [vp, ip]
    lib.tsc_get_layout_stats.argtypes = [vp, C.POINTER(i64)]
    lib.tsc_get_counters.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64)]
    lib.tsc_get_params.argtypes = [vp, dp, dp, dp, dp]
    lib.tsc_set_params.argtypes = [vp, dp, dp]
    lib.tsc_estep.argtypes = [vp, dp, dp, dp]
    lib.tsc_mstep.argtypes = [vp, dp, dp, dp]
    lib.tsc_calculate_lnl.argtypes = [vp, dp, dp, dp, dp]
    lib.tsc_get_z.argtypes = [vp, i32, dp]
    lib.tsc_reassign_nbest.argtypes = [vp, i32, ip]
    lib.tsc_reassign_colsum.argtypes = [vp, i32, dbl, i32, ip, dp]
    lib.tsc_reassign_data.argtypes = [vp, i32, dbl, i32, ip, dp]
    for name in SYMBOLS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("tsc_abi_version",):
            fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc):
    if rc != TSC_OK:
        msg = load().tsc_last_error().decode("utf-8", "replace")
        if rc == TSC_ERR_ARG and msg.startswith('Argument "method"'):
            raise ValueError(msg)
        raise TelescopeCudaError("libtelescope_b200 error %d: %s" % (rc, msg))


def device_count():
    n = C.c_int32(0)
    rc = load().tsc_device_count(C.byref(n))
    return n.value if rc == TSC_OK else 0


def find_nccl():
    """Prefer the NCCL that ships with the CUDA wheels (newer than the system one); None = loader default."""
    env = os.environ.get("TELESCOPE_B200_NCCL")
    if env:
        return env
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        if spec and spec.submodule_search_locations:
            cand = os.path.join(list(spec.submodule_search_locations)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                return cand
    except Exception:
        pass
    return None


def nccl_unique_id():
    lib = load()
    path = find_nccl()
    if path:
        lib.tsc_set_nccl_path(path.encode())
    buf = (C.c_char * 128)()
    check(lib.tsc_nccl_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class PinnedArray(object):
    """numpy array over page-locked host memory (freed with the object)."""

    def __init__(self, shape, dtype):
        self._lib = load()
        dtype = np.dtype(dtype)
        n = int(np.prod(shape))
        self._ptr = self._lib.tsc_pinned_alloc(max(1, n * dtype.itemsize))
        if not self._ptr:
            raise TelescopeCudaError(self._lib.tsc_last_error().decode())
        buf = (C.c_char * (n * dtype.itemsize)).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)

    def __del__(self):
        try:
            if self._ptr:
                self.array = None
                self._lib.tsc_pinned_free(self._ptr)
                self._ptr = None
        except Exception:
            pass
