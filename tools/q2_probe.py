"""Seconds of the two back-transformation stages for k eigenvectors at order n (bk_debug_twostage)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np  # noqa: E402

import krls_oracle as o  # noqa: E402
from bigkrls_b200 import _lib  # noqa: E402
from bigkrls_b200._lib import check, dptr  # noqa: E402

n = int(sys.argv[1])
lib = _lib.load()
ctx = _lib.default_context(0)
X, y = o.synthetic(n, 10, 1003)
Xs, *_ = o.standardize(X, y)
A = np.asfortranarray(o.gauss_kernel(Xs, 10))
d = np.zeros(n)
e = np.zeros(n)
for k in [int(v) for v in sys.argv[2:]]:
    Z = np.asfortranarray(np.random.default_rng(1).standard_normal((n, k)))
    t = (C.c_double * 4)()
    for rep in range(2):
        check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, t))
    print(f"n={n} k={k}: sy2sb {t[0]:.4f} sb2st {t[1]:.4f} q2 {t[2]:.4f} q1 {t[3]:.4f}", flush=True)
