"""Two-stage tridiagonalisation (sy2sb.cu / sb2st.cu) against numpy/LAPACK on the host (GPU).

Checks each stage separately: the stage-1 band has the spectrum of K, the stage-2 tridiagonal has the
spectrum of K, and the back-transformed eigenvectors diagonalise K.  Tolerances: eigenvalues 1e-12 relative
to the largest (the product tolerance is 1e-9), residual / orthogonality 1e-12.
"""
import numpy as np
import pytest
from scipy.linalg import eigh_tridiagonal

import krls_oracle as o
from bigkrls_b200 import _lib
from bigkrls_b200._lib import check, dptr

pytestmark = pytest.mark.gpu


def kernel_matrix(n, p, seed=5):
    X, y = o.synthetic(n, p, seed)
    Xs, *_ = o.standardize(X, y)
    return np.asfortranarray(o.gauss_kernel(Xs, p))


def band_to_dense(band, n, b):
    B = np.zeros((n, n))
    for d in range(b + 1):
        idx = np.arange(n - d)
        B[idx + d, idx] = band[d, idx]
        B[idx, idx + d] = band[d, idx]
    return B


@pytest.mark.parametrize("n,p,k", [(3, 3, 3), (66, 3, 66), (130, 3, 130), (517, 6, 100), (1000, 4, 1000),
                                    (3001, 5, 400), (3001, 5, 36), (4608, 5, 72), (2500, 4, 17)])
def test_twostage(ctx, n, p, k):
    lib = _lib.load()
    A = kernel_matrix(n, p)
    band = np.zeros((128, n), order="F")
    d = np.zeros(n)
    e = np.zeros(n)
    ref = np.linalg.eigvalsh(A)
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, dptr(band), dptr(d), dptr(e), None, 0, None))
    # stage 1
    assert np.all(band[65:, :] == 0.0)
    Bd = band_to_dense(band, n, 64)
    assert np.max(np.abs(np.linalg.eigvalsh(Bd) - ref)) < 1e-12 * ref.max()
    # stage 2
    lam, S = eigh_tridiagonal(d, e[:n - 1]) if n > 1 else (d.copy(), np.ones((1, 1)))
    assert np.max(np.abs(lam - ref)) < 1e-12 * ref.max()
    # back-transformation of the top-k eigenvectors
    Z = np.asfortranarray(S[:, n - k:])
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, None))
    resid = np.max(np.abs(A @ Z - Z * lam[n - k:]))
    orth = np.max(np.abs(Z.T @ Z - np.eye(k)))
    assert resid < 1e-12 * ref.max() * np.sqrt(n), resid
    assert orth < 1e-12 * np.sqrt(n), orth


@pytest.mark.parametrize("n,pipe_min", [(1000, 256), (1531, 512), (3001, 512), (3001, 2048), (4608, 1024)])
def test_sy2sb_delayed_update_pipeline(ctx, monkeypatch, n, pipe_min):
    """The pipelined phase of the dense->band stage (chain of pair i on the matrix BEFORE the previous pair's update,
    corrected with skinny products; updates out of place on the other stream) forced at small n: tridiagonal spectrum
    against LAPACK, and the back-transformed vectors diagonalise K (they go through the panels the phase deposits)."""
    lib = _lib.load()
    A = kernel_matrix(n, 5)
    ref = np.linalg.eigvalsh(A)
    monkeypatch.setenv("BK_SY2SB_PIPE_MIN", str(pipe_min))
    d = np.zeros(n)
    e = np.zeros(n)
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), None, 0, None))
    lam, S = eigh_tridiagonal(d, e[:n - 1])
    assert np.max(np.abs(lam - ref)) < 1e-12 * ref.max()
    k = 64
    Z = np.asfortranarray(S[:, n - k:])
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, None))
    assert np.max(np.abs(A @ Z - Z * lam[n - k:])) < 1e-12 * ref.max() * np.sqrt(n)
    assert np.max(np.abs(Z.T @ Z - np.eye(k))) < 1e-12 * np.sqrt(n)
    # and against the in-place loop alone
    monkeypatch.setenv("BK_SY2SB_NOPIPE", "1")
    d0 = np.zeros(n)
    e0 = np.zeros(n)
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d0), dptr(e0), None, 0, None))
    lam0 = eigh_tridiagonal(d0, e0[:n - 1], eigvals_only=True)
    assert np.max(np.abs(lam - lam0)) < 1e-12 * ref.max()


@pytest.mark.parametrize("n", [130, 193, 200, 257, 517, 1000, 3001, 6007])
def test_chase_handoff_protocols_agree_bitwise(ctx, monkeypatch, n):
    """The early hand-off kernel (tagged slots + staged blocks, sb2st.cu) runs the same arithmetic as the
    completion-flag kernel: tridiagonal, band and back-transformed vectors must be bit-identical (sizes around the
    first full hop, 1 + 3*64 rows, and several CTAs' worth of sweeps)."""
    lib = _lib.load()
    A = kernel_matrix(n, 4)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BK_CHASE_LL", flag)
        d = np.zeros(n)
        e = np.zeros(n)
        check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), None, 0, None))
        k = min(n, 96)
        Z = np.asfortranarray(np.random.default_rng(7).standard_normal((n, k)))
        check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, None))
        out[flag] = (d.copy(), e.copy(), Z.copy())
    assert np.array_equal(out["0"][0], out["1"][0])
    assert np.array_equal(out["0"][1], out["1"][1])
    assert np.array_equal(out["0"][2], out["1"][2])


@pytest.mark.parametrize("n,p,k", [(66, 3, 66), (517, 6, 100), (1000, 4, 1000), (3001, 5, 400)])
def test_q2_blocked_backtransform(ctx, monkeypatch, n, p, k):
    """The GEMM-based (compact-WY blocks) Q2 back-transformation, forced for every k (odd n: unaligned path)."""
    monkeypatch.setenv("BK_Q2_BLOCKED", "1")
    lib = _lib.load()
    A = kernel_matrix(n, p)
    d = np.zeros(n)
    e = np.zeros(n)
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), None, 0, None))
    lam, S = eigh_tridiagonal(d, e[:n - 1])
    Z = np.asfortranarray(S[:, n - k:])
    check(lib.bk_debug_twostage(ctx.handle, dptr(A), n, None, dptr(d), dptr(e), dptr(Z), k, None))
    assert np.max(np.abs(A @ Z - Z * lam[n - k:])) < 1e-12 * lam.max() * np.sqrt(n)
    assert np.max(np.abs(Z.T @ Z - np.eye(k))) < 1e-12 * np.sqrt(n)


def test_twostage_matches_onestage_fit(ctx, monkeypatch):
    """The product path gives the same fit with either tridiagonalisation (tolerances of the north star)."""
    from bigkrls_b200 import api
    X, y = o.synthetic(4500, 4, 11)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BK_EIG_TWOSTAGE", flag)
        fit = api.bigKRLS(y, X, eigtrunc=0.001, ctx=ctx, noisy=False)
        out[flag] = (fit["lambda"], np.array(fit["coeffs"]).ravel().copy(), np.array(fit["derivatives"]).copy(),
                     fit["lastkeeper"])
        fit.release_device()
    assert out["0"][3] == out["1"][3]
    assert abs(out["0"][0] - out["1"][0]) <= 1e-9 * abs(out["0"][0])
    assert np.max(np.abs(out["0"][1] - out["1"][1])) <= 1e-8 * np.max(np.abs(out["0"][1]))
    assert np.max(np.abs(out["0"][2] - out["1"][2])) <= 1e-8 * np.max(np.abs(out["0"][2]))


def test_wide_spectrum_twostage_with_blocked_q2(ctx, monkeypatch):
    """Flat spectrum (many dimensions): the threshold keeps most eigenvectors; up to n = 16384 the two-stage path
    keeps going with the GEMM-based Q2 and must agree with the one-stage path within the north-star tolerances."""
    from bigkrls_b200 import api
    X, y = o.synthetic(4200, 40, 3)
    monkeypatch.delenv("BK_EIG_TWOSTAGE", raising=False)
    monkeypatch.delenv("BK_Q2_BLOCKED", raising=False)
    fit = api.bigKRLS(y, X, eigtrunc=0.0001, ctx=ctx, noisy=False, derivative=False, vcov_est=False)
    assert fit["lastkeeper"] > 4200 // 3
    assert fit["_info"]["twostage"] == 1.0
    c2, lam2, k2 = np.array(fit["coeffs"]).ravel().copy(), fit["lambda"], fit["lastkeeper"]
    fit.release_device()
    monkeypatch.setenv("BK_EIG_TWOSTAGE", "0")
    fit0 = api.bigKRLS(y, X, eigtrunc=0.0001, ctx=ctx, noisy=False, derivative=False, vcov_est=False)
    c1 = np.array(fit0["coeffs"]).ravel()
    assert fit0["lastkeeper"] == k2
    assert abs(fit0["lambda"] - lam2) <= 1e-9 * abs(lam2)
    assert np.max(np.abs(c1 - c2)) <= 1e-8 * np.max(np.abs(c1))
    fit0.release_device()


def test_wide_spectrum_falls_back_to_onestage_at_large_n(ctx, monkeypatch):
    """Above n = 16384 a threshold that keeps more than n/3 eigenvectors sends the solver back to the one-stage
    path (its back-transformation is a plain GEMM): the result is exactly the one-stage result."""
    from bigkrls_b200 import api
    n = 16500
    X, y = o.synthetic(n, 40, 3)
    monkeypatch.delenv("BK_EIG_TWOSTAGE", raising=False)
    fit = api.bigKRLS(y, X, eigtrunc=0.0001, ctx=ctx, noisy=False, derivative=False, vcov_est=False,
                      return_squares=False)
    assert fit["lastkeeper"] > n // 3
    assert fit["_info"]["twostage"] == 0.0
    ref_c, ref_lam = np.array(fit["coeffs"]).copy(), fit["lambda"]
    fit.release_device()
    monkeypatch.setenv("BK_EIG_TWOSTAGE", "0")
    fit0 = api.bigKRLS(y, X, eigtrunc=0.0001, ctx=ctx, noisy=False, derivative=False, vcov_est=False,
                       return_squares=False)
    assert fit0["lambda"] == ref_lam
    assert np.array_equal(np.array(fit0["coeffs"]), ref_c)
    fit0.release_device()
