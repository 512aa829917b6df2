"""Host-side logic of the library (pure CPU): speculative golden-section search against the
sequential oracle, and the D&C deflation bookkeeping against the numpy prototype."""
import ctypes as C

import numpy as np
import pytest

import dc_prototype as dc
import krls_oracle as o
from bigkrls_b200 import _lib
from util import mtcars


def run_host_search(values, n, Q, ys, batch):
    lib = _lib.load()
    calls = []

    def cb(user, lams, nlam, out):
        calls.append(nlam)
        for i in range(nlam):
            out[i] = o.solve_for_c(Q, values, ys, lams[i])[0]
        return 0

    cbf = _lib.LE_CALLBACK(cb)
    lam, L, U = C.c_double(), C.c_double(), C.c_double()
    probes, passes = C.c_int(), C.c_int()
    ev = np.ascontiguousarray(values)
    _lib.check(lib.bk_host_lambda_search(_lib.dptr(ev), len(ev), n, -1.0, 0.0, 0.0, batch,
                                         C.cast(cbf, C.c_void_p), None, C.byref(lam), C.byref(L),
                                         C.byref(U), C.byref(probes), C.byref(passes)))
    return lam.value, L.value, U.value, probes.value, passes.value


@pytest.mark.parametrize("batch", [1, 3, 7, 15])
def test_lambda_search_matches_sequential_oracle(batch):
    for (X, y, trunc) in [mtcars()[1:][::-1] + (0.0,), o.synthetic(300, 4, 11) + (0.0,),
                          o.synthetic(400, 3, 12) + (0.001,)]:
        X, y = (X, y) if X.ndim == 2 else (y, X)
        Xs, ys, *_ = o.standardize(X, y)
        K = o.gauss_kernel(Xs, X.shape[1])
        eo = o.eigen(K, None, trunc)
        ref_lam, ref_probes = o.lambda_search(eo["vectors"], eo["values"], ys)
        L0, U0 = o.lambda_bounds(eo["values"], len(ys))
        lam, L, U, probes, passes = run_host_search(eo["values"], len(ys), eo["vectors"], ys, batch)
        assert (L, U) == (L0, U0)
        assert lam == ref_lam            # bit-identical path
        assert probes == ref_probes
        if batch == 1:
            assert passes == probes
        if batch >= 7:
            assert passes <= (probes + 2) // 3 + 1


def host_deflate_c(d, z, n1, beta):
    lib = _lib.load()
    n = d.size
    K, nrot = C.c_int(), C.c_int()
    dlam, w, dv = np.zeros(n), np.zeros(n), np.zeros(n)
    ndc, ndt, dc_ = (np.zeros(n, np.int32) for _ in range(3))
    ridx, rcs = np.zeros(2 * n, np.int32), np.zeros(2 * n)
    i32 = lambda a: a.ctypes.data_as(_lib.c_int32_p)
    _lib.check(lib.bk_host_deflate_test(_lib.dptr(d), _lib.dptr(z), n, n1, beta, C.byref(K),
                                        _lib.dptr(dlam), _lib.dptr(w), i32(ndc), i32(ndt), i32(dc_),
                                        _lib.dptr(dv), C.byref(nrot), i32(ridx), _lib.dptr(rcs)))
    k, r = K.value, nrot.value
    return dict(K=k, dlam=dlam[:k], w=w[:k], nd_cols=ndc[:k], nd_type=ndt[:k], defl_cols=dc_[:n - k],
                defl_vals=dv[:n - k], rots=[(ridx[2 * i], ridx[2 * i + 1], rcs[2 * i], rcs[2 * i + 1])
                                            for i in range(r)])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_host_deflate_matches_prototype(seed):
    rng = np.random.default_rng(seed)
    n, n1 = 60, 29
    d = np.sort(rng.standard_normal(n))
    d[10:14] = d[10]                      # exact ties -> rotations
    d[40] = d[41] + 1e-18
    z = rng.standard_normal(n)
    z[5] = 1e-20                          # tiny weight -> deflation
    z[33] = 0.0
    d = rng.permutation(d)
    beta = 0.37 * (-1) ** seed
    a = dc.host_deflate(d, z, abs(beta), n1)
    b = host_deflate_c(d, z, n1, beta)
    assert a["K"] == b["K"] and a["K"] < n
    for k in ("dlam", "w", "defl_vals"):
        assert np.array_equal(a[k], b[k]), k
    for k in ("nd_cols", "nd_type", "defl_cols"):
        assert a[k].tolist() == b[k].tolist(), k
    assert len(a["rots"]) == len(b["rots"]) > 0
    for ra, rb in zip(a["rots"], b["rots"]):
        assert tuple(ra) == tuple(rb)


def test_dc_prototype_is_a_valid_eigensolver():
    from scipy.linalg import eigh_tridiagonal
    rng = np.random.default_rng(5)
    n = 257
    d, e = rng.standard_normal(n), rng.standard_normal(n - 1)
    lam, Q, _ = dc.stedc(d, e, leaf=32)
    ref = eigh_tridiagonal(d, e, eigvals_only=True)
    T = np.diag(d) + np.diag(e, 1) + np.diag(e, -1)
    assert np.max(np.abs(lam - ref)) < 1e-13 * np.max(np.abs(ref)) * 10
    assert np.max(np.abs(Q.T @ Q - np.eye(n))) < 1e-13
    assert np.max(np.abs(T @ Q - Q * lam)) < 1e-13 * np.max(np.abs(ref)) * 10


def host_bounds(values, n):
    lib = _lib.load()

    def cb(user, lams, nlam, out):
        for i in range(nlam):
            out[i] = 1.0          # flat loss: the golden section stops at once, only the bounds matter
        return 0

    cbf = _lib.LE_CALLBACK(cb)
    lam, L, U = C.c_double(), C.c_double(), C.c_double()
    probes, passes = C.c_int(), C.c_int()
    ev = np.ascontiguousarray(values, dtype=np.float64)
    rc = lib.bk_host_lambda_search(_lib.dptr(ev), len(ev), n, -1.0, 0.0, 0.0, 7, C.cast(cbf, C.c_void_p), None,
                                   C.byref(lam), C.byref(L), C.byref(U), C.byref(probes), C.byref(passes))
    return rc, L.value, U.value


@pytest.mark.parametrize("seed", range(4))
def test_lambda_bounds_with_signed_noise_tail(seed):
    """A real kernel spectrum ends in rounding noise of both signs; the bounds code sums that tail through suffix sums
    (e/(e+x) ~ e/x) and must still stop exactly where the reference's scans stop."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(1500, 6000))
    lead = int(rng.integers(40, 400))
    ev = np.concatenate([n * np.exp(-np.arange(lead) * (8.0 / lead)), 1e-12 * n * (rng.random(n - lead) - 0.4)])
    ev = np.sort(ev)[::-1]
    ev[:lead] *= (n - ev[lead:].sum()) / ev[:lead].sum()
    L0, U0 = o.lambda_bounds(ev, n)
    rc, L, U = host_bounds(ev, n)
    assert rc == 0 and (L, U) == (L0, U0)


@pytest.mark.parametrize("seed", range(6))
def test_lambda_bounds_bisection_equals_linear_scan(seed):
    """The upper bound is found by gallop + bisection instead of the reference's `U <- U - 1` scan
    (R/bigKRLS_Rcpp_functions.R:16-20); it must stop at exactly the same U for any spectrum."""
    rng = np.random.default_rng(seed)
    n = int(rng.integers(200, 3000))
    shape = seed % 3
    if shape == 0:      # fast geometric decay (Gaussian kernel, few dimensions)
        ev = n * 0.5 ** np.arange(n) + 1e-14 * rng.random(n)
    elif shape == 1:    # one dominant value over a flat floor (many dimensions)
        ev = np.concatenate([[0.15 * n], 0.85 + 0.1 * rng.random(n - 1)])
    else:               # power law
        ev = n / (1.0 + np.arange(n)) ** 1.5
    ev = np.sort(ev)[::-1] * (n / ev.sum())          # trace = n like a unit-diagonal kernel
    L0, U0 = o.lambda_bounds(ev, n)
    rc, L, U = host_bounds(ev, n)
    assert rc == 0 and (L, U) == (L0, U0)


def test_lambda_bounds_degenerate_spectrum_is_an_error():
    rc, _, _ = host_bounds(np.zeros(50), 50)           # sum(ev/(ev+U)) < 1 for every U: the scan never ends
    assert rc != 0


def test_user_supplied_bounds_are_respected_including_zero():
    """R accepts L = 0 (`L >= 0`, R/bigKRLS.R:225-228): a supplied bound is used as is, only a missing one (L < 0 /
    U <= 0 at the C boundary) triggers the reference's bounds loop."""
    lib = _lib.load()
    rng = np.random.default_rng(5)
    n = 400
    ev = np.sort(n / (1.0 + np.arange(n)) ** 1.5)[::-1]
    ev *= n / ev.sum()

    def cb(user, lams, nlam, out):
        for i in range(nlam):
            out[i] = (lams[i] - 3.0) ** 2 * 100.0      # convex, minimum at 3
        return 0

    cbf = _lib.LE_CALLBACK(cb)
    out = {}
    for name, (L, U) in {"both": (0.0, 10.0), "auto": (-1.0, 0.0)}.items():
        lam, Lo, Uo = C.c_double(), C.c_double(), C.c_double()
        probes, passes = C.c_int(), C.c_int()
        evc = np.ascontiguousarray(ev)
        _lib.check(lib.bk_host_lambda_search(_lib.dptr(evc), n, n, L, U, 0.0, 7, C.cast(cbf, C.c_void_p), None,
                                             C.byref(lam), C.byref(Lo), C.byref(Uo), C.byref(probes), C.byref(passes)))
        out[name] = (lam.value, Lo.value, Uo.value)
    assert out["both"][1:] == (0.0, 10.0) and abs(out["both"][0] - 3.0) < 0.5
    L0, U0 = o.lambda_bounds(ev, n)
    assert out["auto"][1:] == (L0, U0)
