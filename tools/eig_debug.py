"""Isolate eigensolver stages on the GPU: determinism + accuracy of sytrd and stedc."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import numpy as np
from scipy.linalg import eigh_tridiagonal
import krls_oracle as o
from bigkrls_b200 import _lib
from bigkrls_b200._lib import check, dptr, fmat

lib = _lib.load(); ctx = _lib.default_context(0)
for (n, p, seed) in [(3100, 5, 6), (600, 3, 5), (2000, 10, 3)]:
    X, y = o.synthetic(n, p, seed)
    Xs, ys, *_ = o.standardize(X, y)
    K = o.gauss_kernel(Xs, p)
    refK = np.linalg.eigvalsh(K)
    ds, es = [], []
    for rep in range(3):
        d, e = np.empty(n), np.empty(n - 1)
        check(lib.bk_debug_sytrd(ctx.handle, dptr(K), n, dptr(d), dptr(e)))
        ds.append(d); es.append(e)
    print(f"n={n} p={p}: sytrd run-to-run max|dd|={max(np.max(np.abs(ds[0]-ds[i])) for i in (1,2)):.3e} "
          f"max|de|={max(np.max(np.abs(es[0]-es[i])) for i in (1,2)):.3e}")
    evT = eigh_tridiagonal(ds[0], es[0], eigvals_only=True)
    err = np.abs(evT - refK)
    print(f"   eig(T_gpu) vs eig(K): max abs err {err.max():.3e} (norm {refK.max():.3e}); worst idx {np.argsort(err)[-3:]}, errs {np.sort(err)[-3:]}")
    evs = []
    for rep in range(2):
        ev, Z = np.empty(n), np.empty((n, n), order="F")
        check(lib.bk_debug_stedc(ctx.handle, dptr(ds[0]), dptr(es[0]), n, dptr(ev), dptr(Z)))
        evs.append(ev)
    err2 = np.abs(evs[0] - evT)
    print(f"   stedc vs scipy on same (d,e): max abs err {err2.max():.3e}; worst idx {np.argsort(err2)[-3:]} errs {np.sort(err2)[-3:]}; "
          f"run-to-run {np.max(np.abs(evs[0]-evs[1])):.3e}")
    T = np.diag(ds[0]) + np.diag(es[0], 1) + np.diag(es[0], -1)
    Zd = Z[:, ::-1]
    print(f"   stedc orth {np.max(np.abs(Zd.T@Zd-np.eye(n))):.3e} resid {np.max(np.abs(T@Zd-Zd*evs[1])):.3e}")
    np.save(f"gpurun_out/dbg_d_{n}.npy", ds[0]); np.save(f"gpurun_out/dbg_e_{n}.npy", es[0])
