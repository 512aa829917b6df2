"""ctypes wrapper of oracle/libkrls_port.so (the compiled literal restatement).  TEST
INFRASTRUCTURE / CPU BASELINE ONLY - see krls_port.cpp."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(os.path.join(_HERE, "libkrls_port.so"))
        dp = C.POINTER(C.c_double)
        _lib.krls_port_fit.restype = C.c_int
        _lib.krls_port_fit.argtypes = [dp, dp, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                       dp, dp, C.POINTER(C.c_int), dp, dp, dp, dp, dp, dp]
    return _lib


def fit(Xs, ys, sigma=None, eigtrunc=0.0, threads=0, derivative=True):
    """Xs, ys standardised.  Returns dict with times (s per stage) and the main outputs."""
    Xs = np.asfortranarray(Xs, dtype=np.float64)
    ys = np.ascontiguousarray(ys, dtype=np.float64)
    n, p = Xs.shape
    sigma = float(p) if sigma is None else float(sigma)
    dp = C.POINTER(C.c_double)
    P = lambda a: a.ctypes.data_as(dp)
    times = np.zeros(8)
    lam, Le, lk = C.c_double(), C.c_double(), C.c_int()
    ev, c, yf = np.empty(n), np.empty(n), np.empty(n)
    D, var = np.empty((n, p), order="F"), np.empty(p)
    rc = lib().krls_port_fit(P(Xs), P(ys), n, p, sigma, float(eigtrunc), int(threads), int(derivative),
                             P(times), C.byref(lam), C.byref(lk), P(ev), P(c), P(yf), P(D), P(var), C.byref(Le))
    if rc != 0:
        raise RuntimeError(f"krls_port_fit failed: {rc}")
    names = ["kernel", "eigen", "lambda", "coef", "vcov", "deriv", "total"]
    return {"times": dict(zip(names, times[:7])), "probes": int(times[7]), "lambda": lam.value,
            "lastkeeper": lk.value, "Le": Le.value, "evals": ev, "coeffs": c, "yfitted_std": yf,
            "derivatives_std": D, "var_std": var, "threads": lib().krls_port_threads()}
