"""Stage-level parity of the CUDA path against the oracle, through the C ABI (GPU)."""
import ctypes as C

import numpy as np
import pytest

import krls_oracle as o
from bigkrls_b200 import _lib
from bigkrls_b200._lib import check, dptr, fmat
from util import corolla_golden, mtcars, relerr

pytestmark = pytest.mark.gpu


def gauss(ctx, X, sigma):
    X = fmat(X)
    n, p = X.shape
    K = np.empty((n, n), order="F")
    check(_lib.load().bk_gauss_kernel(ctx.handle, dptr(X), n, p, float(sigma), dptr(K)))
    return K


def dgemm(ctx, ta, tb, A, B):
    A, B = fmat(A), fmat(B)
    m = A.shape[1] if ta else A.shape[0]
    k = A.shape[0] if ta else A.shape[1]
    n = B.shape[0] if tb else B.shape[1]
    Cm = np.empty((m, n), order="F")
    check(_lib.load().bk_dgemm(ctx.handle, int(ta), int(tb), m, n, k, dptr(A), A.shape[0], dptr(B),
                               B.shape[0], dptr(Cm), m))
    return Cm


@pytest.mark.parametrize("n,p", [(1, 1), (63, 3), (64, 16), (300, 5), (1037, 17), (2500, 5)])
def test_gauss_kernel(ctx, n, p):
    rng = np.random.default_rng(n + p)
    X = rng.standard_normal((n, p))
    K = gauss(ctx, X, p)
    ref = o.gauss_kernel(X, p)
    assert np.max(np.abs(K - ref)) < 1e-14          # tolerance: 1e-14 absolute (values in (0, 1])
    assert np.array_equal(K, K.T)                    # exactly symmetric
    assert np.all(np.diag(K) == 1.0)                 # exact ones on the diagonal, like the reference


def test_gauss_kernel_mtcars_golden(ctx):
    # reference tests/testthat/test_basic_usage.R:62-99
    names, y, X = mtcars()
    Xs, *_ = o.standardize(X, y)
    K = gauss(ctx, Xs, X.shape[1])
    g = corolla_golden()
    j = names.index("Toyota Corolla")
    assert max(abs(K[i, j] - g[nm]) for i, nm in enumerate(names)) < 1e-13


@pytest.mark.parametrize("m,n,p", [(5, 7, 2), (200, 333, 10), (1000, 129, 20)])
def test_temp_kernel(ctx, m, n, p):
    rng = np.random.default_rng(m)
    A, B = fmat(rng.standard_normal((m, p))), fmat(rng.standard_normal((n, p)))
    out = np.empty((m, n), order="F")
    check(_lib.load().bk_temp_kernel(ctx.handle, dptr(A), m, dptr(B), n, p, float(p), dptr(out)))
    assert np.max(np.abs(out - o.temp_kernel(A, B, p))) < 1e-14


@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
@pytest.mark.parametrize("m,n,k", [(1, 1, 1), (7, 5, 3), (128, 128, 16), (129, 257, 65), (300, 22, 1000),
                                   (64, 64, 5000), (513, 40, 77), (33, 700, 130)])
def test_dgemm(ctx, ta, tb, m, n, k):
    rng = np.random.default_rng(m * 7 + n * 3 + k)
    A = rng.standard_normal((k, m) if ta else (m, k))
    B = rng.standard_normal((n, k) if tb else (k, n))
    got = dgemm(ctx, ta, tb, A, B)
    ref = (A.T if ta else A) @ (B.T if tb else B)
    assert relerr(got, ref) < 1e-13                  # FP64 FMA accumulation, k <= 5000


def test_crossprod_family(ctx):
    lib = _lib.load()
    rng = np.random.default_rng(3)
    A, B = fmat(rng.standard_normal((211, 13))), fmat(rng.standard_normal((211, 9)))
    out = np.empty((13, 9), order="F")
    check(lib.bk_crossprod(ctx.handle, dptr(A), 211, 13, dptr(B), 9, dptr(out)))
    assert relerr(out, A.T @ B) < 1e-13
    out = np.empty((13, 13), order="F")
    check(lib.bk_xtx(ctx.handle, dptr(A), 211, 13, dptr(out)))
    assert relerr(out, A.T @ A) < 1e-13
    B2 = fmat(rng.standard_normal((77, 13)))
    out = np.empty((211, 77), order="F")
    check(lib.bk_tcrossprod(ctx.handle, dptr(A), 211, 13, dptr(B2), 77, dptr(out)))
    assert relerr(out, A @ B2.T) < 1e-13
    out = np.empty((211, 211), order="F")
    check(lib.bk_xxt(ctx.handle, dptr(A), 211, 13, dptr(out)))
    assert relerr(out, A @ A.T) < 1e-13
    d = rng.standard_normal(13)
    out = np.empty((211, 13), order="F")
    check(lib.bk_mult_diag(ctx.handle, dptr(A), 211, 13, dptr(d), dptr(out)))
    assert np.array_equal(out, A * d[None, :])


def _eig_setup(n, p, seed, trunc):
    X, y = o.synthetic(n, p, seed)
    Xs, ys, *_ = o.standardize(X, y)
    K = o.gauss_kernel(Xs, p)
    eo = o.eigen(K, None, trunc)
    return Xs, ys, K, eo


@pytest.mark.parametrize("n,trunc", [(100, 0.0), (700, 0.0), (1500, 0.001)])
def test_loo_and_solve_for_c(ctx, n, trunc):
    lib = _lib.load()
    Xs, ys, K, eo = _eig_setup(n, 4, n, trunc)
    Q, ev = fmat(eo["vectors"]), np.ascontiguousarray(eo["values"])
    k = Q.shape[1]
    for nl in (1, 2, 3, 7, 15):
        lams = np.ascontiguousarray(np.geomspace(0.01, 50, nl))
        Le = np.empty(nl)
        check(lib.bk_loo_batch(ctx.handle, dptr(Q), n, k, dptr(ev), dptr(ys), dptr(lams), nl, dptr(Le)))
        ref = np.array([o.solve_for_c(Q, ev, ys, l)[0] for l in lams])
        assert relerr(Le, ref) < 1e-10 and np.max(np.abs(Le / ref - 1)) < 1e-9
    Le1, c = C.c_double(), np.empty(n)
    check(lib.bk_solve_for_c(ctx.handle, dptr(Q), n, k, dptr(ev), dptr(ys), 0.37, C.byref(Le1), dptr(c)))
    rLe, rc = o.solve_for_c(Q, ev, ys, 0.37, literal=(n <= 700))
    assert abs(Le1.value / rLe - 1) < 1e-9 and relerr(c, rc) < 1e-10


def _check_eig(A, vals, vecs, tol_val=1e-12):
    n, k = vecs.shape
    ref = np.linalg.eigvalsh(A)[::-1]
    scale = np.max(np.abs(ref))
    # eigenvalues: |d lambda| <= 1e-12 * lambda_max (absolute, relative to the norm), which is
    # <= 1e-9 relative for every eigenvalue >= 1e-3 * lambda_max (BASELINE.json tolerance)
    assert np.max(np.abs(vals - ref[: vals.size])) <= tol_val * scale
    assert np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vecs.T @ vecs - np.eye(k))) < 5e-13
    assert np.max(np.abs(A @ vecs - vecs * vals[:k])) < 5e-12 * scale


@pytest.mark.parametrize("n", [2, 3, 31, 32, 33, 64, 65, 100, 257, 600, 1037])
def test_eigen_full_kernel_matrix(ctx, n):
    lib = _lib.load()
    rng = np.random.default_rng(n)
    X = rng.standard_normal((n, 4))
    A = o.gauss_kernel(X, 4.0)
    vals, vecs = np.empty(n), np.empty((n, n), order="F")
    check(lib.bk_eigen(ctx.handle, dptr(A), n, n, dptr(vals), dptr(vecs)))
    _check_eig(A, vals, vecs)


@pytest.mark.parametrize("kind", ["random", "clustered", "diag", "rank1"])
def test_eigen_full_other_spectra(ctx, kind):
    lib = _lib.load()
    n = 300
    rng = np.random.default_rng(9)
    if kind == "random":
        B = rng.standard_normal((n, n))
        A = (B + B.T) / 2
    elif kind == "clustered":
        Qm, _ = np.linalg.qr(rng.standard_normal((n, n)))
        A = (Qm * np.repeat([1.0, 2.0, 2.0 + 1e-13, 5.0, -3.0, 0.0], n // 6)) @ Qm.T
        A = (A + A.T) / 2
    elif kind == "diag":
        A = np.diag(rng.standard_normal(n))
    else:
        v = rng.standard_normal(n)
        A = np.outer(v, v)
    A = fmat(A)
    vals, vecs = np.empty(n), np.empty((n, n), order="F")
    check(lib.bk_eigen(ctx.handle, dptr(A), n, n, dptr(vals), dptr(vecs)))
    _check_eig(A, vals, vecs, tol_val=2e-12)


def test_eigen_top_neig_and_values_only(ctx):
    lib = _lib.load()
    n, neig = 500, 40
    X = np.random.default_rng(1).standard_normal((n, 3))
    A = o.gauss_kernel(X, 3.0)
    vals, vecs = np.empty(neig), np.empty((n, neig), order="F")
    check(lib.bk_eigen(ctx.handle, dptr(A), n, neig, dptr(vals), dptr(vecs)))
    _check_eig(A, vals, vecs)
    vals2 = np.empty(n)
    check(lib.bk_eigen(ctx.handle, dptr(A), n, n, dptr(vals2), None))
    assert np.array_equal(vals2[:neig], vals)


@pytest.mark.parametrize("n,neig,p", [(1500, 120, 5), (2048, 64, 10), (3000, 500, 20)])
def test_eigen_topk_krylov(ctx, n, neig, p):
    # Neig << N -> restarted block-Krylov path (reference: sp_mat + eigs_sym, src/eigen.cpp:18-22)
    lib = _lib.load()
    X, y = o.synthetic(n, p, 1004)
    Xs, *_ = o.standardize(X, y)
    A = o.gauss_kernel(Xs, p)
    vals, vecs = np.empty(neig), np.empty((n, neig), order="F")
    check(lib.bk_eigen(ctx.handle, dptr(A), n, neig, dptr(vals), dptr(vecs)))
    _check_eig(A, vals, vecs)


@pytest.mark.parametrize("binary", [False, True])
def test_deriv_mat(ctx, binary):
    lib = _lib.load()
    n, p = 400, 5
    X, y = o.synthetic(n, p, 21, binary_last=binary)
    f = o.bigkrls(y, X, eigtrunc=0)
    Xs, ys, *_ = o.standardize(X, y)
    V = fmat(f["vcov.est.c"] / np.std(y, ddof=1) ** 2)
    K, c = fmat(f["K"]), np.ascontiguousarray(f["coeffs"].reshape(-1))
    D, var = np.empty((n, p), order="F"), np.empty(p)
    check(lib.bk_deriv_mat(ctx.handle, dptr(Xs), n, p, dptr(K), dptr(V), dptr(c), float(p), dptr(D), dptr(var)))
    rD, rvar = o.deriv_mat(Xs, K, V, c, float(p), literal=True)
    assert relerr(D, rD) < 1e-10
    assert np.max(np.abs(var / rvar - 1)) < 1e-8


def test_neffective(ctx):
    lib = _lib.load()
    X = fmat(np.random.default_rng(4).standard_normal((777, 6)))
    out = C.c_double()
    check(lib.bk_neffective(ctx.handle, dptr(X), 777, 6, C.byref(out)))
    assert abs(out.value / o.neffective_acf(X) - 1) < 1e-12


def test_error_behaviour(ctx):
    lib = _lib.load()
    X = fmat(np.zeros((4, 2)))
    K = np.empty((4, 4), order="F")
    assert lib.bk_gauss_kernel(ctx.handle, dptr(X), 4, 2, -1.0, dptr(K)) == -1
    assert b"sigma" in lib.bk_last_error()
    assert lib.bk_gauss_kernel(None, dptr(X), 4, 2, 1.0, dptr(K)) == -1
    A = fmat(np.full((8, 8), np.nan))
    vals = np.empty(8)
    assert lib.bk_eigen(ctx.handle, dptr(A), 8, 8, dptr(vals), None) == -3   # BK_ERR_NUMERIC
