import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from scipy.linalg import eigh_tridiagonal
from bigkrls_b200 import _lib
from bigkrls_b200._lib import check, dptr
lib = _lib.load(); ctx = _lib.default_context(0)
rng = np.random.default_rng(0)
for n in [int(a) for a in sys.argv[1:]]:
    X = rng.standard_normal((n, 5))
    G = X @ X.T
    sq = np.diag(G)
    K = np.asfortranarray(np.exp(-(sq[:, None] + sq[None, :] - 2 * G) / 5.0))
    K = (K + K.T) / 2
    ref = np.linalg.eigvalsh(K)
    out = []
    for rep in range(2):
        d, e = np.empty(n), np.empty(n - 1)
        check(lib.bk_debug_sytrd(ctx.handle, dptr(K), n, dptr(d), dptr(e)))
        out.append((d, e))
    evT = eigh_tridiagonal(out[0][0], out[0][1], eigvals_only=True)
    print(f"n={n}: err {np.max(np.abs(evT-ref)):.2e} run2run {np.max(np.abs(out[0][0]-out[1][0])):.2e}", flush=True)
