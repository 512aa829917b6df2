"""End-to-end parity of bigKRLS()/predict()/crossvalidate through the public API (GPU).

Tolerances are BASELINE.json's: eigenvalues and lambda* 1e-9 relative, coefficients, fitted
values and marginal effects 1e-8 relative (max|d|/max|ref| per field)."""
import numpy as np
import pytest

import krls_oracle as o
from bigkrls_b200 import bigKRLS, crossvalidate_bigKRLS, predict, summary
from util import corolla_golden, mtcars, relerr

pytestmark = pytest.mark.gpu

FIELDS_1E8 = ["coeffs", "yfitted", "derivatives", "avgderivatives", "var.avgderivatives", "vcov.est.c",
              "vcov.est.fitted", "K"]


def compare(fit, ref, fields=FIELDS_1E8):
    ev, rev = fit["K.eigenvalues"], ref["K.eigenvalues"]
    big = rev >= 1e-3 * rev[0]
    assert np.max(np.abs(ev[big] / rev[big] - 1)) < 1e-9          # retained eigenvalues: 1e-9 relative
    assert np.max(np.abs(ev - rev)) < 1e-9 * rev[0] * 1e-3        # the rest: 1e-12 * lambda_1 absolute
    assert fit["lastkeeper"] == ref["lastkeeper"], (fit["lastkeeper"], ref["lastkeeper"])
    assert abs(fit["lambda"] / ref["lambda"] - 1) < 1e-9
    for k in fields:
        if k in ref and ref[k] is not None and k in fit:
            assert relerr(fit[k], ref[k]) < 1e-8, k
    for k in ("Looe", "Neffective", "R2", "R2AME"):
        if k in ref:
            assert abs(fit[k] / ref[k] - 1) < 1e-8, k


def test_mtcars_reference_known_answers():
    # reference tests/testthat/test_basic_usage.R:39-100
    names, y, X = mtcars()
    fit = bigKRLS(y, X, eigtrunc=0, Ncores=1)
    g = corolla_golden()
    j = names.index("Toyota Corolla")
    assert max(abs(fit["K"][i, j] - g[nm]) for i, nm in enumerate(names)) < 0.01     # :99 (actual ~1e-15)
    assert max(abs(fit["K"][i, j] - g[nm]) for i, nm in enumerate(names)) < 1e-13
    Xn = X.copy()
    Xn[:, 2] = 200
    fc = predict(fit, Xn)
    assert np.mean(fc["predicted"] < y) == 0.6875                                     # :58
    compare(fit, o.bigkrls(y, X, eigtrunc=0, literal=True))
    assert fit["binaryindicator"].tolist() == [False] * 6 + [True, True] + [False] * 2
    s = summary(fit)
    assert s["ttests"].shape == (10, 4) and s["rownames"][6].endswith("*")
    # "bigmemory example works" (:103-123): Neig = nrow(X)
    fit2 = bigKRLS(y, X, Neig=X.shape[0])
    assert relerr(fit2["yfitted"], fit["yfitted"]) < 1e-12


@pytest.mark.parametrize("n,p,seed,kw", [
    (500, 5, 1, dict(eigtrunc=0)),
    (1200, 6, 2, dict(eigtrunc=0.001)),
    (800, 4, 3, dict(eigtrunc=0, binary=True)),
    (900, 8, 4, dict(eigtrunc=0.01, which_derivatives=[1, 3, 5])),
    (600, 5, 5, dict(lambda_=0.5)),
    (3100, 5, 6, dict()),                       # n > 3000 -> default eigtrunc = 0.001
    (2400, 8, 7, dict(Neig=300, eigtrunc=0.001, which_derivatives=[1, 3, 5])),   # config-4 shape: Krylov path
])
def test_fit_parity_synthetic(n, p, seed, kw):
    kw = dict(kw)
    binary = kw.pop("binary", False)
    X, y = o.synthetic(n, p, seed, binary_last=binary)
    okw = {("lam" if k == "lambda_" else k): v for k, v in kw.items()}
    fit = bigKRLS(y, X, **kw)
    ref = o.bigkrls(y, X, **okw)
    if fit["lastkeeper"] != ref["lastkeeper"]:
        # Only legitimate when eigtrunc == 0 and the trailing eigenvalues are rounding noise: the
        # reference's lastkeeper = max(which(values >= 0)) then depends on the sign of that noise.
        ev = ref["K.eigenvalues"]
        lo, hi = sorted((fit["lastkeeper"], ref["lastkeeper"]))
        assert kw.get("eigtrunc", 1) == 0 and np.max(np.abs(ev[lo - 1:hi])) < 1e-13 * ev[0]
        ref = o.bigkrls(y, X, force_lastkeeper=fit["lastkeeper"], **okw)
    compare(fit, ref)
    assert fit["_info"]["n_probes"] == ref["_nprobe"]
    if binary:
        assert fit["binaryindicator"][-1]


def test_config1_shape_n2500_p5():
    # BASELINE.json configs[0]: N=2500 P=5, all derivatives
    X, y = o.synthetic(2500, 5, 1001)
    ref = o.bigkrls(y, X)
    fit = bigKRLS(y, X)
    compare(fit, ref)


def test_predict_and_se():
    X, y = o.synthetic(700, 4, 9)
    Xn, _ = o.synthetic(150, 4, 10)
    ref = o.bigkrls(y, X, eigtrunc=0)
    fit = bigKRLS(y, X, eigtrunc=0)
    rp = o.predict(ref, Xn, se_pred=True)
    gp = predict(fit, Xn, se_pred=True)
    assert relerr(gp["predicted"], rp["predicted"]) < 1e-8
    assert relerr(gp["newdataK"], rp["newdataK"]) < 1e-13
    assert relerr(gp["se.pred"], rp["se.pred"]) < 1e-7


def test_crossvalidate_folds():
    X, y = o.synthetic(600, 4, 1005)
    folds = np.random.default_rng(1005).permutation(600) % 3 + 1
    ref = o.crossvalidate_folds(y, X, folds)
    got = crossvalidate_bigKRLS(y, X, folds=folds)
    for k, v in ref.items():
        assert relerr(got[k], v) < 1e-7, k


def test_validation_errors_match_reference_messages():
    X, y = o.synthetic(50, 3, 1)
    with pytest.raises(ValueError, match="eigtrunc must be between 0"):
        bigKRLS(y, X, eigtrunc=2)
    Xc = X.copy()
    Xc[:, 1] = 3.0
    with pytest.raises(ValueError, match="are constant and must be removed: 2"):
        bigKRLS(y, Xc)
    with pytest.raises(ValueError, match="y is a constant"):
        bigKRLS(np.ones(50), X)
    with pytest.raises(ValueError, match="vcov.est is needed"):
        bigKRLS(y, X, vcov_est=False)
    with pytest.raises(ValueError, match="nrow\\(X\\) not equal"):
        bigKRLS(y[:-1], X)
    f = bigKRLS(y, X, derivative=False, vcov_est=False)
    assert "derivatives" not in f and "vcov.est.c" not in f


@pytest.mark.parametrize("n,p,trunc", [(1100, 3, 0.01), (2300, 4, 0.001), (3000, 6, 0.001)])
def test_dc_factored_top_level_agrees(monkeypatch, n, p, trunc):
    """Divide & conquer with the root's children kept factored (few eigenvectors wanted: the two largest merge GEMMs
    are skipped, stedc.cu) against the plain tree: same eigenvalues, lambda and lastkeeper, coefficients and variances
    at rounding level - and against the oracle at the usual tolerances."""
    X, y = o.synthetic(n, p, 77)
    out = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("BK_DC_LAZY", flag)
        fit = bigKRLS(y, X, eigtrunc=trunc, noisy=False)
        out[flag] = (fit["lambda"], int(fit["lastkeeper"]), np.array(fit["K.eigenvalues"]).copy(),
                     np.array(fit["coeffs"]).ravel().copy(), np.array(fit["vcov.est.c"]).copy(),
                     np.array(fit["var.avgderivatives"]).ravel().copy())
        fit.release_device()
    a, b = out["0"], out["1"]
    assert a[1] == b[1] and abs(a[0] / b[0] - 1) < 1e-9
    # the root's z comes from rows of the children's eigenvector matrices, formed by a GEMM in one variant and by a
    # row-vector product in the other: eigenvalues agree at rounding level, not bit for bit
    assert np.max(np.abs(a[2] - b[2])) <= 1e-13 * a[2][0]
    assert relerr(b[3], a[3]) < 1e-10 and relerr(b[4], a[4]) < 1e-10 and relerr(b[5], a[5]) < 1e-10
    ref = o.bigkrls(y, X, eigtrunc=trunc)
    assert b[1] == ref["lastkeeper"] and abs(b[0] / ref["lambda"] - 1) < 1e-9
    assert relerr(b[3], np.asarray(ref["coeffs"]).ravel()) < 1e-8
