#!/usr/bin/env python
"""Generates the full-size parity fixtures under tests/golden/ with the CPU oracle.

    python tools/make_fixtures.py [c1 c2 c3 c4r c5 ge]          (default: all)

Runs on CPU only (numpy + scipy's LAPACK dsyevd = the routine arma::eig_sym calls, src/eigen.cpp:24) with
`oracle.bigkrls(literal=False)` on the SURVEY.md 8d synthetic workloads at the sizes BASELINE.json states:

    c2   N=10 000 P=10 eigtrunc=0 (all eigenvectors)              seed 1002
    c3   N=20 000 P=10 eigtrunc=0.001 (the headline config)       seed 1003
    c4r  N=20 000 P=20 Neig=500 which.derivatives=c(1,3,5)        seed 1004   (config 4 at the largest N whose
         dense eigh fits this container; the full N=60 000 needs 29 GB per matrix and hours of LAPACK)
    c5   5-fold cross-validation at N=20 000 P=10, fold vector seed 1005
    ge   examples/data2016GE.csv of the reference (3106 x 68, 50 binary columns): BigDerivMat at scale

Each fixture is a compact .npz: every eigenvalue, lambda*, probe count, lastkeeper, coefficients, fitted values,
derivatives, average derivatives and their variances, the scalar fit statistics, and a fixed 256 x 256 scattered
sample (rows ridx, columns cidx) of each N x N field.  The -m gpu tests (tests/test_gpu_golden.py) compare the CUDA
path with these at the north_star tolerances; the GPU box has no /root/reference and needs none.
"""
import gzip
import os
import shutil
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import krls_oracle as o  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
NS = 256


def sample_idx(n, seed):
    rng = np.random.default_rng(seed)
    return np.sort(rng.choice(n, size=min(NS, n), replace=False)), np.sort(rng.choice(n, size=min(NS, n), replace=False))


def pack_fit(ref, ridx, cidx, extra=None):
    d = {"evals": ref["K.eigenvalues"], "lambda": ref["lambda"], "nprobe": ref["_nprobe"],
         "lastkeeper": ref["lastkeeper"], "Le": ref["_Le"], "Looe": ref["Looe"], "Neffective": ref["Neffective"],
         "R2": ref["R2"], "coeffs": ref["coeffs"].reshape(-1), "yfitted": ref["yfitted"],
         "ridx": ridx, "cidx": cidx, "K_blk": ref["K"][np.ix_(ridx, cidx)]}
    if "vcov.est.c" in ref:
        d["Vc_blk"] = ref["vcov.est.c"][np.ix_(ridx, cidx)]
        d["Vf_blk"] = ref["vcov.est.fitted"][np.ix_(ridx, cidx)]
        d["sigmasq"] = ref["sigmasq"]
    if "derivatives" in ref:
        d["derivatives"] = ref["derivatives"]
        d["avgderivatives"] = ref["avgderivatives"].reshape(-1)
        d["var_avgderivatives"] = ref["var.avgderivatives"].reshape(-1)
        d["R2AME"] = ref["R2AME"]
    if extra:
        d.update(extra)
    return d


def save(name, d):
    path = os.path.join(GOLD, name + ".npz")
    np.savez_compressed(path, **d)
    print("wrote %s (%.2f MB)" % (path, os.path.getsize(path) / 1e6), flush=True)


def fit_fixture(name, N, P, seed, **kw):
    t0 = time.time()
    X, y = o.synthetic(N, P, seed)
    ref = o.bigkrls(y, X, **kw)
    ridx, cidx = sample_idx(N, seed + 7)
    # a prediction on 64 held-out points of the same generator (predict.bigKRLS, R/bigKRLS.R:590-621)
    Xn, _ = o.synthetic(64, P, seed + 1)
    pr = o.predict(ref, Xn, se_pred=True)
    extra = {"N": N, "P": P, "seed": seed, "pred_X": Xn, "pred": pr["predicted"], "pred_se": pr["se.pred"].reshape(-1),
             "pred_vcov": pr["vcov.est.pred"]}
    save(name, pack_fit(ref, ridx, cidx, extra))
    print("%s: lambda %.9f lastkeeper %d probes %d (%.0f s)" % (name, ref["lambda"], ref["lastkeeper"], ref["_nprobe"],
                                                               time.time() - t0), flush=True)


def c5_fixture():
    N, P, seed = 20000, 10, 1005
    X, y = o.synthetic(N, P, seed)
    folds = np.random.default_rng(seed).permutation(N) % 5 + 1      # stored beside the data
    stats = {k: [] for k in ("R2_is", "R2_oos", "MSE_is", "MSE_oos", "R2AME_is", "R2AME_oos", "MSE_AME_is", "MSE_AME_oos",
                             "lambda", "lastkeeper", "nprobe")}
    for k in range(1, 6):
        t0 = time.time()
        tr, te = folds != k, folds == k
        trained = o.bigkrls(y[tr], X[tr])
        tested = o.predict(trained, X[te])
        delta = trained["avgderivatives"].reshape(-1)
        yhat = X[te] @ delta
        stats["R2_is"].append(trained["R2"])
        stats["R2_oos"].append(np.corrcoef(y[te], tested["predicted"])[0, 1] ** 2)
        stats["MSE_is"].append(np.mean((y[tr] - trained["yfitted"]) ** 2))
        stats["MSE_oos"].append(np.mean((y[te] - tested["predicted"]) ** 2))
        stats["R2AME_is"].append(trained["R2AME"])
        stats["MSE_AME_is"].append(np.mean((y[tr] - X[tr] @ delta) ** 2))
        stats["R2AME_oos"].append(np.corrcoef(y[te], yhat)[0, 1] ** 2)
        stats["MSE_AME_oos"].append(np.mean((y[te] - yhat) ** 2))
        stats["lambda"].append(trained["lambda"])
        stats["lastkeeper"].append(trained["lastkeeper"])
        stats["nprobe"].append(trained["_nprobe"])
        print("c5 fold %d: lambda %.9f lastkeeper %d (%.0f s)" % (k, trained["lambda"], trained["lastkeeper"],
                                                                   time.time() - t0), flush=True)
    d = {k: np.array(v) for k, v in stats.items()}
    d.update({"N": N, "P": P, "seed": seed, "folds": folds.astype(np.int8)})
    save("c5_cv_N20000_P10", d)


def ge_fixture():
    """The reference's election data set (3106 counties x 68 columns, 50 state dummies): the binary branch of
    src/bigderiv_v3.cpp:31-87 at scale.  y = column 1 (gop_2016_delta), X = the other 67 (examples/cv_election2016.R)."""
    src = "/root/reference/examples/data2016GE.csv"
    dst = os.path.join(GOLD, "data2016GE.csv.gz")
    if not os.path.exists(dst):
        with open(src, "rb") as f, gzip.open(dst, "wb", compresslevel=9) as g:
            shutil.copyfileobj(f, g)
    raw = np.loadtxt(gzip.open(dst, "rt"), delimiter=",", skiprows=1)
    y, X = raw[:, 0], np.asfortranarray(raw[:, 1:])
    t0 = time.time()
    ref = o.bigkrls(y, X)                   # N = 3106 > 3000 -> eigtrunc 0.001 (R/bigKRLS.R:195-201)
    ridx, cidx = sample_idx(X.shape[0], 2016)
    save("ge2016_fit", pack_fit(ref, ridx, cidx, {"binary": ref["binaryindicator"]}))
    print("ge: lambda %.9f lastkeeper %d binary columns %d (%.0f s)" % (ref["lambda"], ref["lastkeeper"],
                                                                        int(ref["binaryindicator"].sum()), time.time() - t0),
          flush=True)


JOBS = {
    "c1": lambda: fit_fixture("c1_N2500_P5", 2500, 5, 1001),
    "c2": lambda: fit_fixture("c2_N10000_P10", 10000, 10, 1002, eigtrunc=0.0),
    "c3": lambda: fit_fixture("c3_N20000_P10", 20000, 10, 1003, eigtrunc=0.001),
    "c4r": lambda: fit_fixture("c4r_N20000_P20_Neig500", 20000, 20, 1004, Neig=500, which_derivatives=[1, 3, 5]),
    "c5": c5_fixture,
    "ge": ge_fixture,
}

if __name__ == "__main__":
    for j in (sys.argv[1:] or list(JOBS)):
        JOBS[j]()
