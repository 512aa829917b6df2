"""ctypes binding of libbigkrls_b200.so (the C ABI declared in include/bigkrls_b200.h).

The library is the product; this module only marshals numpy buffers into it.  There is no
Python/numpy fallback: if the shared library is missing or no B200 is visible every compute
call raises.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbigkrls_b200.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class BKError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"bigkrls_b200 error {status}: {message}")
        self.status = status
        self.message = message


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)
ALLGATHERV_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, c_int64_p, c_int64_p)
BROADCAST_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int)
LE_CALLBACK = C.CFUNCTYPE(C.c_int, C.c_void_p, c_double_p, C.c_int, c_double_p)
EXCHANGE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64)


class Comm(C.Structure):
    _fields_ = [("rank", C.c_int), ("world", C.c_int), ("user", C.c_void_p),
                ("allreduce_sum", ALLREDUCE_FN), ("allgatherv", ALLGATHERV_FN),
                ("broadcast", BROADCAST_FN), ("peer", C.c_void_p)]


class FitOpts(C.Structure):
    _fields_ = [("sigma", C.c_double), ("eigtrunc", C.c_double), ("neig", C.c_int64),
                ("lambda_", C.c_double), ("L", C.c_double), ("U", C.c_double), ("tol", C.c_double),
                ("derivative", C.c_int), ("vcov", C.c_int), ("n_which", C.c_int),
                ("which", c_int32_p), ("y_sd", C.c_double), ("loo_batch", C.c_int),
                ("keep_vcov_fitted", C.c_int), ("K_host", C.c_void_p)]


class FitInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("p", C.c_int64), ("neig", C.c_int64), ("lastkeeper", C.c_int64),
                ("n_deriv", C.c_int64),
                ("lambda_", C.c_double), ("Le", C.c_double), ("sigmasq", C.c_double),
                ("neffective", C.c_double),
                ("n_probes", C.c_int), ("n_passes", C.c_int),
                ("t_kernel", C.c_double), ("t_eigen", C.c_double), ("t_lambda", C.c_double),
                ("t_coef", C.c_double), ("t_vcov", C.c_double), ("t_deriv", C.c_double),
                ("t_total", C.c_double),
                ("t_tridiag", C.c_double), ("t_dc", C.c_double), ("t_backtransform", C.c_double),
                ("sytrd_launches", C.c_double), ("sytrd_kernel_seconds", C.c_double),
                ("sytrd_bytes", C.c_double),
                ("dc_levels", C.c_double), ("dc_merge_flops", C.c_double), ("dc_top_n", C.c_double),
                ("dc_top_k", C.c_double), ("gpu_launches", C.c_double),
                ("krylov_matvecs", C.c_double), ("krylov_restarts", C.c_double),
                ("twostage", C.c_double), ("t_sy2sb", C.c_double), ("t_sb2st", C.c_double),
                ("t_q2", C.c_double), ("t_q1", C.c_double),
                ("band_gemm_launches", C.c_double), ("band_gemm_seconds", C.c_double),
                ("band_gemm_flops", C.c_double)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            d[name.rstrip("_")] = getattr(self, name)
        return d


# name -> (restype, argtypes); mirrors include/bigkrls_b200.h one to one
_SIGNATURES = {
    "bk_version": (C.c_int, []),
    "bk_last_error": (C.c_char_p, []),
    "bk_init": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "bk_destroy": (None, [C.c_void_p]),
    "bk_device_info": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_int), c_int64_p, c_int64_p]),
    "bk_launch_count": (C.c_int64, [C.c_void_p]),
    "bk_trim": (C.c_int, [C.c_void_p]),
    "bk_host_alloc": (C.c_int, [C.c_void_p, C.c_int64, C.POINTER(C.c_void_p)]),
    "bk_host_free": (C.c_int, [C.c_void_p, C.c_void_p]),
    "bk_gauss_kernel": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, C.c_double, c_double_p]),
    "bk_temp_kernel": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, C.c_int64, C.c_int64,
                                 C.c_double, c_double_p]),
    "bk_eigen": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p]),
    "bk_solve_for_c": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p,
                                 C.c_double, c_double_p, c_double_p]),
    "bk_loo_batch": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p,
                               c_double_p, C.c_int, c_double_p]),
    "bk_mult_diag": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p]),
    "bk_crossprod": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, c_double_p]),
    "bk_xtx": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "bk_tcrossprod": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, C.c_int64, c_double_p]),
    "bk_xxt": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "bk_deriv_mat": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p, c_double_p,
                               c_double_p, C.c_double, c_double_p, c_double_p]),
    "bk_neffective": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, C.c_int64, c_double_p]),
    "bk_dgemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, c_double_p,
                           C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_int64]),
    "bk_peer_create": (C.c_int, [C.c_void_p, C.c_int, C.c_int, EXCHANGE_FN, C.c_void_p, C.POINTER(C.c_void_p)]),
    "bk_peer_destroy": (None, [C.c_void_p]),
    "bk_peer_selftest": (C.c_int, [C.c_void_p, C.POINTER(C.c_int)]),
    "bk_fit_default_opts": (None, [C.POINTER(FitOpts), C.c_int64, C.c_int64]),
    "bk_fit_run": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, C.c_int64, C.POINTER(FitOpts),
                             C.POINTER(Comm), C.POINTER(C.c_void_p)]),
    "bk_fit_run_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64,
                                    C.POINTER(FitOpts), C.POINTER(Comm), C.POINTER(C.c_void_p)]),
    "bk_fit_free": (None, [C.c_void_p]),
    "bk_fit_get_info": (C.c_int, [C.c_void_p, C.POINTER(FitInfo)]),
    "bk_fit_col_range": (C.c_int, [C.c_void_p, c_int64_p, c_int64_p]),
    "bk_fit_get_K": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_eigenvalues": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_eigenvectors": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_coeffs": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_yfitted": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_vcov_c": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_vcov_fitted": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_derivatives": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_var_avgderiv": (C.c_int, [C.c_void_p, c_double_p]),
    "bk_fit_get_binary": (C.c_int, [C.c_void_p, c_int32_p]),
    "bk_fit_predict": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p]),
    "bk_fit_predict_full": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p,
                                      c_double_p]),
    "bk_microbench": (C.c_int, [C.c_void_p, C.c_int, C.c_int64, C.c_int, c_double_p]),
    "bk_dgemm_bench": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int,
                                 C.c_double, C.c_int, c_double_p]),
    "bk_debug_gemm": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_double, c_double_p,
                                C.c_int64, c_double_p, C.c_int64, C.c_double, c_double_p, C.c_int64, C.c_int, C.c_int]),
    "bk_debug_sytrd": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "bk_debug_twostage": (C.c_int, [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p,
                                    c_double_p, C.c_int64, c_double_p]),
    "bk_debug_stedc": (C.c_int, [C.c_void_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]),
    "bk_host_lambda_search": (C.c_int, [c_double_p, C.c_int64, C.c_int64, C.c_double, C.c_double, C.c_double,
                                        C.c_int, C.c_void_p, C.c_void_p, c_double_p, c_double_p, c_double_p,
                                        C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "bk_host_deflate_test": (C.c_int, [c_double_p, c_double_p, C.c_int, C.c_int, C.c_double,
                                       C.POINTER(C.c_int), c_double_p, c_double_p, c_int32_p, c_int32_p,
                                       c_int32_p, c_double_p, C.POINTER(C.c_int), c_int32_p, c_double_p]),
}

_lib = None


def load():
    """Load the shared library (raises if it has not been built: python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise BKError(-2, f"{LIB_PATH} not found - build it with `python -c 'import __graft_entry__ as g; "
                          f"g.build()'` (nvcc, sm_100a); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise BKError(status, load().bk_last_error().decode("utf-8", "replace"))


def dptr(a):
    """numpy float64 array -> double* (array must be F- or C-contiguous and stay alive)."""
    return a.ctypes.data_as(c_double_p)


def fmat(a):
    """column-major float64 copy/view of a 2-D array (R / bigmemory layout)."""
    a = np.asarray(a, dtype=np.float64)
    if a.ndim == 1:
        a = a.reshape(-1, 1)
    return np.asfortranarray(a)


class Context:
    """One library context per process x device."""

    def __init__(self, device=0):
        self._lib = load()
        h = C.c_void_p()
        check(self._lib.bk_init(int(device), C.byref(h)))
        self.handle = h
        self.device = int(device)

    def close(self):
        if getattr(self, "handle", None):
            self._lib.bk_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def trim(self):
        """Give cached device memory back (eigensolver work matrices kept between fits, unused pool blocks)."""
        check(self._lib.bk_trim(self.handle))

    def device_info(self):
        name = C.create_string_buffer(256)
        sm = C.c_int()
        tot = C.c_int64()
        free = C.c_int64()
        check(self._lib.bk_device_info(self.handle, name, 256, C.byref(sm), C.byref(tot), C.byref(free)))
        return {"name": name.value.decode(), "sm_count": sm.value, "hbm_total": tot.value,
                "hbm_free": free.value}

    def pinned_empty(self, shape, order="F"):
        """float64 array in pinned (page-locked) host memory.  Buffers come from a per-context
        pool: recycle_pinned(arr) returns one for reuse (cudaHostAlloc of GBs is slow)."""
        n = int(np.prod(shape))
        nbytes = max(8, 8 * n)
        pool = self.__dict__.setdefault("_pool", {})
        if pool.get(nbytes):
            addr = pool[nbytes].pop()
        else:
            p = C.c_void_p()
            check(self._lib.bk_host_alloc(self.handle, nbytes, C.byref(p)))
            addr = p.value
        buf = (C.c_double * n).from_address(addr)
        arr = np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape, order=order)
        self.__dict__.setdefault("_live", {})[addr] = nbytes
        return arr

    def recycle_pinned(self, arr):
        """Give a pinned_empty() array back to the pool (the caller must drop its references)."""
        addr = arr.ctypes.data
        live = self.__dict__.get("_live", {})
        if addr in live:
            self.__dict__.setdefault("_pool", {}).setdefault(live.pop(addr), []).append(addr)

    def drain_pinned(self):
        for lst in self.__dict__.get("_pool", {}).values():
            while lst:
                self._lib.bk_host_free(self.handle, C.c_void_p(lst.pop()))


_default_ctx = {}


def default_context(device=None):
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0")) if "BIGKRLS_DEVICE" not in os.environ \
            else int(os.environ["BIGKRLS_DEVICE"])
    if device not in _default_ctx:
        _default_ctx[device] = Context(device)
    return _default_ctx[device]
