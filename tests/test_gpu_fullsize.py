"""BASELINE.json's full size (N = 20 000, P = 10, eigtrunc = 0.001, all derivatives) checked through
size-independent properties - the oracle's O(N^3) work does not finish in seconds at this size (GPU)."""
import numpy as np
import pytest

from bigkrls_b200 import bigKRLS
from bigkrls_b200 import _lib
import ctypes as C

pytestmark = pytest.mark.gpu


def test_config3_properties():
    rng = np.random.default_rng(1003)
    N, P = 20000, 10
    X = rng.standard_normal((N, P))
    y = np.sin(X[:, 0]) + X[:, 1] * X[:, 2] + 0.5 * rng.standard_normal(N)
    fit = bigKRLS(y, X, eigtrunc=0.001)
    ev, k, lam = fit["K.eigenvalues"], fit["lastkeeper"], fit["lambda"]
    K = fit["K"]
    # checksum of checksums: trace(K) = N exactly (unit diagonal) and ||K||_F^2 = sum of squared eigenvalues
    assert abs(ev.sum() / N - 1) < 1e-10
    assert abs(np.sum(ev ** 2) / np.sum(K * K) - 1) < 1e-10
    assert np.all(np.diff(ev) <= 0) and ev[k - 1] >= 1e-3 * ev[0] > ev[k]
    # eigenpairs: orthonormal, small residual (relative to lambda_1)
    Q = np.empty((N, k), order="F")
    _lib.check(_lib.load().bk_fit_get_eigenvectors(fit._fit, _lib.dptr(Q)))
    assert np.max(np.abs(Q.T @ Q - np.eye(k))) < 1e-12
    assert np.max(np.abs(K @ Q - Q * ev[:k])) < 1e-11 * ev[0]
    # coefficients and fitted values recomputed on the host from (Q, ev, lambda)
    ys = (y - y.mean()) / np.std(y, ddof=1)
    c = Q @ ((Q.T @ ys) / (ev[:k] + lam))
    assert np.max(np.abs(fit["coeffs"].reshape(-1) - c)) < 1e-8 * np.max(np.abs(c))
    yhat = (K @ c) * np.std(y, ddof=1) + y.mean()
    assert np.max(np.abs(fit["yfitted"] - yhat)) < 1e-8 * np.max(np.abs(yhat))
    # vcov.c is symmetric and equals sigma^2 Q (ev+lam)^-2 Q' on a sampled block
    V = fit["vcov.est.c"]
    assert np.array_equal(V[:500, :500], V[:500, :500].T)
    s2 = fit["_info"]["sigmasq"] * np.var(y, ddof=1)
    blk = (Q[:300] * (s2 / (ev[:k] + lam) ** 2)) @ Q[:400].T
    assert np.max(np.abs(V[:300, :400] - blk)) < 1e-8 * np.max(np.abs(blk))
    # marginal effects: linearity in c  (D is linear in the coefficient vector)
    Xs = (X - X.mean(0)) / np.std(X, axis=0, ddof=1)
    j = 3
    D = (-2.0 / P) * (Xs[:2000, j] * (K[:2000] @ c) - K[:2000] @ (Xs[:, j] * c))
    D = D * np.std(y, ddof=1) / np.std(X[:, j], ddof=1)
    assert np.max(np.abs(fit["derivatives"][:2000, j] - D)) < 1e-8 * np.max(np.abs(D))
    fit.release_device()
