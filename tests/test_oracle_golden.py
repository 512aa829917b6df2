"""The oracle against the reference's own known answers (CPU)."""
import numpy as np

import krls_oracle as o
from util import corolla_golden, mtcars, relerr


def test_kernel_column_golden():
    # reference tests/testthat/test_basic_usage.R:62-99 (tolerance there: 0.01)
    names, y, X = mtcars()
    fit = o.bigkrls(y, X, eigtrunc=0)
    g = corolla_golden()
    j = names.index("Toyota Corolla")
    diff = max(abs(fit["K"][i, j] - g[nm]) for i, nm in enumerate(names))
    assert diff < 1e-13


def test_predict_known_answer():
    # reference tests/testthat/test_basic_usage.R:55-58: mean(predicted < mpg) == 0.6875
    names, y, X = mtcars()
    fit = o.bigkrls(y, X, eigtrunc=0)
    Xn = X.copy()
    Xn[:, 2] = 200  # hp
    pr = o.predict(fit, Xn)
    assert np.mean(pr["predicted"] < y) == 0.6875
    assert abs(fit["lambda"] - 0.1305908251510043) < 1e-12
    assert fit["binaryindicator"].tolist() == [False] * 6 + [True, True] + [False] * 2


def test_literal_equals_reduced_forms():
    # the O(N^3) literal restatement (reference structure) == the reduced forms
    names, y, X = mtcars()
    a = o.bigkrls(y, X, eigtrunc=0, literal=True)
    b = o.bigkrls(y, X, eigtrunc=0, literal=False)
    for k in ("coeffs", "yfitted", "vcov.est.c", "vcov.est.fitted", "derivatives", "var.avgderivatives",
              "avgderivatives"):
        assert relerr(b[k], a[k]) < 1e-12, k
    assert a["lambda"] == b["lambda"]


def test_literal_equals_reduced_synthetic_binary():
    X, y = o.synthetic(150, 4, 7, binary_last=True)
    a = o.bigkrls(y, X, literal=True)
    b = o.bigkrls(y, X, literal=False)
    assert b["binaryindicator"][-1]
    for k in ("coeffs", "derivatives", "var.avgderivatives", "vcov.est.fitted"):
        assert relerr(b[k], a[k]) < 1e-11, k


def test_internal_identities():
    X, y = o.synthetic(200, 3, 3)
    f = o.bigkrls(y, X, eigtrunc=0)
    Xs, ys, *_ = o.standardize(X, y)
    K = f["K"]
    # (K + lambda I) c = y when nothing is truncated
    assert relerr((K + f["lambda"] * np.eye(200)) @ f["coeffs"].reshape(-1), ys) < 1e-8
    # which.derivatives quirk B.1: column i divided by X.init.sd[i]
    g = o.bigkrls(y, X, eigtrunc=0, which_derivatives=[1, 3])
    sd = o.col_sd(X)
    assert relerr(g["derivatives"][:, 1] * sd[1], f["derivatives"][:, 2] * sd[2]) < 1e-12
