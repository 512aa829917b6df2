"""Small two-stage run for compute-sanitizer (memcheck / racecheck): band reduction, chase, both Q2 variants, Q1."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np  # noqa: E402
from scipy.linalg import eigh_tridiagonal  # noqa: E402

import krls_oracle as o  # noqa: E402
from bigkrls_b200 import _lib  # noqa: E402
from bigkrls_b200._lib import check, dptr  # noqa: E402

n, p, k = (int(sys.argv[1]), 4, int(sys.argv[2])) if len(sys.argv) > 2 else (331, 4, 90)
lib = _lib.load()
ctx = _lib.default_context(0)
X, y = o.synthetic(n, p, 5)
Xs, *_ = o.standardize(X, y)
A = np.asfortranarray(o.gauss_kernel(Xs, p))
d, e = np.zeros(n), np.zeros(n)
check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), None, 0, None))
lam, S = eigh_tridiagonal(d, e[:n - 1])
for blocked in (os.environ.get("ORDER", "0,1").split(",")):
    os.environ["BK_Q2_BLOCKED"] = blocked
    Z = np.array(S[:, n - k:], order="F", copy=True)
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, None))
    print("blocked", blocked, "residual", np.max(np.abs(A @ Z - Z * lam[n - k:])))
