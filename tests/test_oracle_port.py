"""The compiled literal port (CPU baseline) against the numpy oracle and the golden fit (CPU)."""
import numpy as np

import krls_oracle as o
import port
from util import mtcars, relerr


def _check(X, y, eigtrunc):
    Xs, ys, xm, xs, ym, ysd = o.standardize(X, y)
    r = port.fit(Xs, ys, eigtrunc=eigtrunc, threads=2)
    ref = o.bigkrls(y, X, eigtrunc=eigtrunc, literal=True)
    assert r["lastkeeper"] == ref["lastkeeper"]
    assert abs(r["lambda"] / ref["lambda"] - 1) < 1e-10
    assert r["probes"] == ref["_nprobe"]
    assert relerr(r["coeffs"], ref["coeffs"].reshape(-1)) < 1e-9
    assert relerr(r["yfitted_std"] * ysd + ym, ref["yfitted"]) < 1e-10
    D = ysd * r["derivatives_std"] / o.col_sd(X)[None, :]
    assert relerr(D, ref["derivatives"]) < 1e-9
    assert relerr((ysd / o.col_sd(X)) ** 2 * r["var_std"], ref["var.avgderivatives"].reshape(-1)) < 1e-8


def test_port_mtcars():
    _, y, X = mtcars()
    _check(X, y, 0.0)


def test_port_synthetic_with_binary_column():
    X, y = o.synthetic(220, 4, 5, binary_last=True)
    _check(X, y, 0.001)
