"""CPU oracle for the bigKRLS estimation hot path (TEST INFRASTRUCTURE ONLY).

This file is a numpy restatement of the reference's algorithm.  It is the checker the
CUDA path is compared against; it is never imported by the product package
(`bigkrls_b200/`).  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.

Parity status: PINNED by the reference's own two known answers
(`tests/testthat/test_basic_usage.R:55-58` -> 0.6875 and `:62-99` -> the 32-value kernel
column), see `tests/test_oracle_golden.py`.  No reference test pins eigenvalues, lambda,
coefficients, vcov or derivatives numerically; for those fields the
oracle is validated by internal identities (literal O(N^3) form == reduced form).

Every function cites the reference file:line it follows (paths relative to the reference
checkout).  Two variants exist where the reference's algorithmic structure matters:

* ``literal=True``  - the same loop/GEMM structure as the C++/R reference (used for small
  N and as the timed "reference CPU path");
* ``literal=False`` - algebraically identical reduced forms, used
  for parity at sizes where the O(N^3)-per-column literal form is too slow.

All arrays are float64; matrices are stored Fortran-order (column-major) like R.
"""
from __future__ import annotations

import numpy as np

EPS = float(np.finfo(np.float64).eps)  # .Machine$double.eps (R/bigKRLS_Rcpp_functions.R:28), exactly 2^-52
GOLD = 0.381966     # R/bigKRLS_Rcpp_functions.R:38-39


# --------------------------------------------------------------------------------------
# A.1 pre-processing  (R/bigKRLS.R:179, 242, 245-254)
# --------------------------------------------------------------------------------------
def col_sd(X):
    """biganalytics::colsd -> sample sd with n-1 denominator (R/bigKRLS.R:179)."""
    return np.std(np.asarray(X, dtype=np.float64), axis=0, ddof=1)


def standardize(X, y):
    """R/bigKRLS.R:245-254: x <- (x - mean)/sd (n-1), same for y."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    xm, xs = X.mean(axis=0), col_sd(X)
    ym, ys = y.mean(), np.std(y, ddof=1)
    Xs = np.asfortranarray((X - xm) / xs)
    return Xs, (y - ym) / ys, xm, xs, ym, ys


def binary_indicator(X):
    """R/bigKRLS.R:242: apply(X, 2, function(x) length(unique(x))) == 2."""
    X = np.asarray(X)
    return np.array([np.unique(X[:, j]).size == 2 for j in range(X.shape[1])])


# --------------------------------------------------------------------------------------
# a1 / a2 Gaussian kernels (src/gauss_kernel.cpp:18-23, src/temp_kernel.cpp:19-23)
# --------------------------------------------------------------------------------------
def gauss_kernel(X, sigma):
    """K_ij = exp(-sum_d (x_id - x_jd)^2 / sigma); direct differences like the reference."""
    X = np.asarray(X, dtype=np.float64)
    n = X.shape[0]
    K = np.empty((n, n), order="F")
    blk = max(1, min(n, int(2e7 // max(1, n * X.shape[1]))))
    for i0 in range(0, n, blk):
        d = X[i0:i0 + blk, None, :] - X[None, :, :]
        K[i0:i0 + blk, :] = np.exp(-1 * np.sum(d * d, axis=2) / sigma)
    return K


def temp_kernel(A, B, sigma):
    """out[i,j] = exp(-||a_i - b_j||^2 / sigma)   (src/temp_kernel.cpp:19-23)."""
    A = np.asarray(A, dtype=np.float64)
    B = np.asarray(B, dtype=np.float64)
    out = np.empty((A.shape[0], B.shape[0]), order="F")
    blk = max(1, min(A.shape[0], int(2e7 // max(1, B.shape[0] * B.shape[1]))))
    for i0 in range(0, A.shape[0], blk):
        d = A[i0:i0 + blk, None, :] - B[None, :, :]
        out[i0:i0 + blk, :] = np.exp(-1 * np.sum(d * d, axis=2) / sigma)
    return out


# --------------------------------------------------------------------------------------
# a3 eigen (src/eigen.cpp:18-29, R/bigKRLS_Rcpp_functions.R:173-199)
# --------------------------------------------------------------------------------------
def eigen(K, Neig=None, eigtrunc=0.0, force_lastkeeper=None):
    """arma::eig_sym (LAPACK dsyevd) then flip to descending (src/eigen.cpp:24,28-29).

    For Neig < N the reference calls arma::eigs_sym (largest magnitude, tol = eps): the
    converged answer equals the top-Neig pairs of the full decomposition, which is what is
    returned here.  R side: vecs <- -1*vecs (:186), lastkeeper (:190), keep cols 1..lastkeeper
    (:192-196) but ALL Neig values.
    """
    import scipy.linalg as sla
    n = K.shape[0]
    Neig = n if Neig is None else int(min(n, Neig))
    vals, vecs = sla.eigh(np.asarray(K), driver="evd")
    vals = vals[::-1].copy()
    vecs = vecs[:, ::-1]
    if Neig < n:
        # eigs_sym default "lm" (largest magnitude); a PSD kernel matrix makes that the
        # algebraically largest ones.
        vals, vecs = vals[:Neig], vecs[:, :Neig]
    vecs = -1.0 * vecs
    lastkeeper = int(np.max(np.nonzero(vals >= eigtrunc * vals[0])[0])) + 1
    if force_lastkeeper is not None:
        # test aid: with eigtrunc = 0 and a numerically singular K the reference's rule
        # `max(which(values >= 0))` is decided by the SIGN OF ROUNDING NOISE in eigenvalues of
        # size ~1e-16*lambda_1 (two LAPACK builds disagree); parity is then checked at equal
        # lastkeeper.
        lastkeeper = int(force_lastkeeper)
    return {"values": vals, "vectors": np.asfortranarray(vecs[:, :lastkeeper]),
            "lastkeeper": lastkeeper}


# --------------------------------------------------------------------------------------
# a4 LOO loss + coefficients (src/solveforc.cpp:36-58)
# --------------------------------------------------------------------------------------
def solve_for_c(Q, values, y, lam, literal=False):
    """Returns (Le, coeffs).  Only the first k = ncol(Q) eigenvalues enter."""
    Q = np.asarray(Q)
    n, k = Q.shape
    w = 1.0 / (np.asarray(values)[:k] + lam)
    if literal:
        # row-by-row lower triangle of Ginv = Q diag(w) Q' (src/solveforc.cpp:36-53)
        Qt = np.ascontiguousarray(Q.T)             # :34 Eigenvectors = trans(Eigenvectors)
        ginv_diag = np.zeros(n)
        coeffs = np.zeros(n)
        for i in range(n):
            ginv = (Qt[:, i] * w) @ Qt[:, :i + 1]  # :42
            ginv_diag[i] = ginv[i]                 # :44
            coeffs[:i] += ginv[:i] * y[i]          # :45 (span(0,i-1); empty at i=0)
            coeffs[i] += np.sum(ginv * y[:i + 1])  # :46
        Le = float(np.sum((coeffs / ginv_diag) ** 2))  # :56-58
        return Le, coeffs
    z = Q.T @ y
    coeffs = Q @ (z * w)
    ginv_diag = (Q * Q) @ w
    return float(np.sum((coeffs / ginv_diag) ** 2)), coeffs


# --------------------------------------------------------------------------------------
# a5 lambda search (R/bigKRLS_Rcpp_functions.R:5-82)
# --------------------------------------------------------------------------------------
def lambda_bounds(values, n):
    """U (:16-25) and L (:26-36).  Uses ALL Neig eigenvalues."""
    ev = np.asarray(values, dtype=np.float64)
    U = float(n)
    while np.sum(ev / (ev + U)) < 1:
        U -= 1
    L = EPS
    q = int(np.argmin(np.abs(ev - np.max(ev) / 1000))) + 1   # which.min is 1-based, first minimum
    while np.sum(ev / (ev + L)) > q:
        L += 0.05
    return L, U


def lambda_search(Q, values, y, L=None, U=None, tol=None, literal=False, trace=None):
    """Golden-section search, step for step (R/bigKRLS_Rcpp_functions.R:38-77).

    NB bigKRLS() never forwards `tol` (R/bigKRLS.R:274-275) so the effective tolerance is
    always 1e-3*n (:10-12)."""
    n = len(y)
    if tol is None:
        tol = 10 ** -3 * n
    L0, U0 = (None, None)
    if L is None or U is None:
        L0, U0 = lambda_bounds(values, n)
    U = U0 if U is None else U
    L = L0 if L is None else L
    loo = lambda lam: solve_for_c(Q, values, y, lam, literal=literal)[0]
    X1 = L + GOLD * (U - L)
    X2 = U - GOLD * (U - L)
    S1 = loo(X1)
    S2 = loo(X2)
    nprobe = 2
    if trace is not None:
        trace.append((L, X1, X2, U, S1, S2))
    while abs(S1 - S2) > tol:
        if S1 < S2:
            U = X2
            X2 = X1
            X1 = L + GOLD * (U - L)
            S2 = S1
            S1 = loo(X1)
        else:
            L = X1
            X1 = X2
            X2 = U - GOLD * (U - L)
            S1 = S2
            S2 = loo(X2)
        nprobe += 1
        if trace is not None:
            trace.append((L, X1, X2, U, S1, S2))
    return (X1 if S1 < S2 else X2), nprobe


# --------------------------------------------------------------------------------------
# a9 marginal effects (src/bigderiv_v3.cpp:24-110)
# --------------------------------------------------------------------------------------
def deriv_mat(X, K, V, coeffs, sigma, literal=False):
    """BigDerivMat on the (standardised) columns of X.  Returns (D N x P', Var P').

    literal=True follows the reference's N x N temporaries and triple products;
    literal=False uses the reduced forms of SURVEY Appendix A.5."""
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    D = np.full((n, p), -1.0, order="F")
    var = np.full(p, -1.0)
    c = np.asarray(coeffs).reshape(-1)
    for j in range(p):
        x = X[:, j]
        uniq = np.unique(x)
        if uniq.size == 2:                                   # :31 binary case
            z0, z1 = x.min(), x.max()                        # :34-35
            sdXj = 1 / (z1 - z0)                             # :36
            phi = -1 / (sdXj ** 2 * sigma)                   # :37
            if literal:
                adj_T = np.empty((n, n))
                adj_C = np.empty((n, n))
                KT = np.empty(n)
                KC = np.empty(n)
                for i in range(n):                           # :49-78
                    c1 = 1 if x[i] == z0 else 0
                    both_max = (x + x[i] == 2 * z1).astype(float)
                    both_min = (x + x[i] == 2 * z0).astype(float)
                    first_greater = (x[i] > x).astype(float)
                    second_greater = (x[i] < x).astype(float)
                    adj_T_local = both_min - first_greater
                    adj_C_local = both_max - second_greater
                    adj_T[i, :] = adj_T_local + first_greater - second_greater
                    adj_C[i, :] = adj_C_local - first_greater + second_greater
                    KT[i] = np.exp(adj_T_local * phi) @ K[:, i]
                    KC[i] = np.exp(adj_C_local * phi) @ K[:, i]
                    c2 = np.exp((-2 * (both_max + both_min) + 1) * (z1 - z0) ** 2 / sigma)
                    D[i, j] = (sdXj * (-1) ** c1 * (1 - c2) * K[:, i]) @ c
                MT = np.exp(adj_T * phi) * K
                MC = np.exp(adj_C * phi) * K
                sT = np.sum(MT @ V.T, axis=0)
                sC = np.sum(MC @ V.T, axis=0)
                vcv_sum = np.sum(sT * KT + sC * KC - 2 * sT * KC)   # :82-84
            else:
                b1 = (x == z1).astype(float)
                b0 = 1.0 - b1
                S0, S1 = K.T @ b0, K.T @ b1
                C0, C1 = K.T @ (b0 * c), K.T @ (b1 * c)
                ep, em = np.exp(phi), np.exp(-phi)
                t = np.where(b0 == 1, ep * S0 + S1, em * S0 + S1)
                u = np.where(b0 == 1, S0 + em * S1, S0 + ep * S1)
                D[:, j] = np.where(b0 == 1,
                                   -sdXj * ((1 - ep) * C0 + (1 - em) * C1),
                                   sdXj * ((1 - em) * C0 + (1 - ep) * C1))
                r = t - u
                vcv_sum = r @ (V @ r)
            var[j] = 2 * sdXj ** 2 * vcv_sum / n ** 2        # :85
        else:                                                # :90-106 continuous
            if literal:
                diff = x[:, None] - x[None, :]               # differences.col(i) = X.col(j) - X(i,j)
                Lm = diff * K
                D[:, j] = (-2 / sigma) * (Lm @ c)
                var[j] = (1 / n ** 2) * (-2 / sigma) ** 2 * np.sum(Lm.T @ V @ Lm)
            else:
                K1 = K @ np.ones(n)
                Kx = K @ x
                D[:, j] = (-2 / sigma) * (x * (K @ c) - K @ (x * c))
                r = x * K1 - Kx
                var[j] = (4 / (sigma ** 2 * n ** 2)) * (r @ (V @ r))
    return D, var


# --------------------------------------------------------------------------------------
# a10 Neffective (acf)  (src/Neffective.cpp:23-64)
# --------------------------------------------------------------------------------------
def neffective_acf(X):
    X = np.asarray(X, dtype=np.float64)
    n, p = X.shape
    Z = X - (X.sum(axis=1) / p)[:, None]
    Z = Z / np.sqrt(np.sum(Z ** 2, axis=1))[:, None]
    r = 0.0
    blk = 1024
    for i0 in range(0, n, blk):
        G = np.abs(Z[i0:i0 + blk] @ Z.T)
        rows = np.arange(i0, min(n, i0 + blk))[:, None]
        cols = np.arange(n)[None, :]
        r += float(np.sum(G[cols < rows]))
    return n * (1 - 2 * r / n ** 2) + 1


# --------------------------------------------------------------------------------------
# bigKRLS() driver  (R/bigKRLS.R:97-516)
# --------------------------------------------------------------------------------------
def bigkrls(y, X, sigma=None, derivative=True, which_derivatives=None, vcov_est=True,
            Neig=None, eigtrunc=None, lam=None, L=None, U=None, acf=False, literal=False,
            force_lastkeeper=None):
    """Returns a dict with the reference's output-list field names (R/bigKRLS.R:420-469).

    `which_derivatives` is 1-based like R.  Quirk B.1 (X.init.sd[i] index bug,
    R/bigKRLS.R:395-397) is replicated."""
    X0 = np.asarray(X, dtype=np.float64)
    y0 = np.asarray(y, dtype=np.float64).reshape(-1)
    n, p = X0.shape
    w = {}
    X_init_sd = col_sd(X0)                                               # :179
    acf = bool(acf) and p > 2                                            # :192
    Neig = min(n, int(Neig)) if Neig is not None else n                 # :194
    if eigtrunc is None:
        eigtrunc = 0.001 if n > 3000 else 0.0                            # :195-201
    sigma = float(p) if sigma is None else float(sigma)                  # :230
    x_is_binary = binary_indicator(X0)                                   # :242
    Xs, ys, xm, xs, y_mean, y_sd = standardize(X0, y0)                   # :245-254
    K = gauss_kernel(Xs, sigma)                                          # :262
    eo = eigen(K, Neig, eigtrunc, force_lastkeeper)                      # :266
    w["K.eigenvalues"] = eo["values"]
    w["lastkeeper"] = eo["lastkeeper"]
    nprobe = 0
    if lam is None:                                                      # :270-278
        lam, nprobe = lambda_search(eo["vectors"], eo["values"], ys, L=L, U=U, literal=literal)
    w["Neffective"] = n - np.sum(eo["values"] / (eo["values"] + lam))    # :280
    Le, coeffs = solve_for_c(eo["vectors"], eo["values"], ys, lam, literal=literal)   # :286
    yfitted = K @ coeffs                                                 # :291
    Q, k = eo["vectors"], eo["lastkeeper"]
    if vcov_est:
        sigmasq = float((ys - yfitted) @ (ys - yfitted)) / n             # :294
        m = Q * (sigmasq * (eo["values"][:k] + lam) ** -2)[None, :]      # :299 (multdiag.cpp:17-18)
        vcovmatc = np.asfortranarray(m @ Q.T)                            # :301 (crossprod.cpp:53)
        if literal:
            vcovmatyhat = K.T @ (vcovmatc @ K)                           # :307
        else:
            m2 = Q * (sigmasq * (eo["values"][:k] / (eo["values"][:k] + lam)) ** 2)[None, :]
            vcovmatyhat = m2 @ Q.T
        w["sigmasq"] = sigmasq
    if derivative:
        wd = None if which_derivatives is None else [int(i) - 1 for i in which_derivatives]
        X_est = Xs if wd is None else Xs[:, wd]                          # :321
        Dm, varavg = deriv_mat(X_est, K, vcovmatc, coeffs, sigma, literal=literal)   # :329
        avg = Dm.mean(axis=0)
        yhat_ame = X_est @ avg                                           # :388
        w["R2AME"] = float(np.corrcoef(y0, yhat_ame)[0, 1] ** 2)         # :390
        Dm = y_sd * Dm                                                   # :392
        for i in range(Dm.shape[1]):
            Dm[:, i] = Dm[:, i] / X_init_sd[i]                           # :393-395  (quirk B.1: index i)
        w["avgderivatives"] = Dm.mean(axis=0)[None, :]                   # :398
        sdsel = X_init_sd if wd is None else X_init_sd[wd]
        w["var.avgderivatives"] = ((y_sd / sdsel) ** 2 * varavg)[None, :]   # :401-405
        w["derivatives"] = np.asfortranarray(Dm)
    if acf:
        w["Neffective.acf"] = neffective_acf(Xs)                         # :412-416
    w["coeffs"] = coeffs.reshape(-1, 1)
    w["y"] = y0
    w["sigma"] = sigma
    w["lambda"] = lam
    w["binaryindicator"] = x_is_binary
    w["which.derivatives"] = which_derivatives
    w["yfitted"] = yfitted * y_sd + y_mean                               # :429
    w["R2"] = 1 - np.var(y0 - w["yfitted"], ddof=1) / y_sd ** 2          # :430
    w["Looe"] = Le * y_sd                                                # :431
    w["K"] = K
    w["X"] = X0
    if vcov_est:
        w["vcov.est.c"] = y_sd ** 2 * vcovmatc                           # :439
        w["vcov.est.fitted"] = y_sd ** 2 * vcovmatyhat                   # :446
    w["derivative.call"] = derivative
    w["_nprobe"] = nprobe
    w["_Le"] = Le
    return w


# --------------------------------------------------------------------------------------
# predict.bigKRLS  (R/bigKRLS.R:547-637)
# --------------------------------------------------------------------------------------
def predict(obj, newdata, se_pred=False, correct_SE=True):
    X = np.asarray(obj["X"], dtype=np.float64)
    new = np.asarray(newdata, dtype=np.float64)
    Xmeans, Xsd = X.mean(axis=0), col_sd(X)                              # :587-588
    Xs = (X - Xmeans) / Xsd                                              # :590-591
    news = (new - Xmeans) / Xsd                                          # :593-594
    newK = temp_kernel(news, Xs, obj["sigma"])                           # :596
    ypred = newK @ obj["coeffs"].reshape(-1)                             # :598
    out = {"newdataK": newK}
    y = np.asarray(obj["y"]).reshape(-1)
    if se_pred:
        vy = np.var(y, ddof=1)
        vp = vy * (newK @ (obj["vcov.est.c"] * (1 / vy))) @ newK.T       # :605
        if correct_SE and obj.get("Neffective") is not None:
            vp = np.sqrt(X.shape[0] / obj["Neffective"]) * vp            # :607-608
        out["vcov.est.pred"] = vp
        out["se.pred"] = np.sqrt(np.diag(vp)).reshape(-1, 1)             # :610
    out["predicted"] = ypred * np.std(y, ddof=1) + y.mean()              # :618
    return out


# --------------------------------------------------------------------------------------
# crossvalidate.bigKRLS, K folds, explicit fold vector  (R/bigKRLS.R:1228-1317)
# --------------------------------------------------------------------------------------
def crossvalidate_folds(y, X, folds, **kw):
    """`folds` is the integer fold vector (1..Kfolds) the reference would draw at :1232."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    folds = np.asarray(folds)
    Kf = int(folds.max())
    out = {k: [] for k in ("R2_is", "R2_oos", "MSE_is", "MSE_oos", "R2AME_is", "R2AME_oos",
                           "MSE_AME_is", "MSE_AME_oos")}
    for k in range(1, Kf + 1):
        tr, te = folds != k, folds == k
        trained = bigkrls(y[tr], X[tr], **kw)
        tested = predict(trained, X[te])
        out["R2_is"].append(trained["R2"])                                          # :1293
        out["R2_oos"].append(np.corrcoef(y[te], tested["predicted"])[0, 1] ** 2)   # :1294
        out["MSE_is"].append(np.mean((y[tr] - trained["yfitted"]) ** 2))           # :1295
        out["MSE_oos"].append(np.mean((y[te] - tested["predicted"]) ** 2))         # :1296
        if "avgderivatives" in trained:
            delta = trained["avgderivatives"].reshape(-1)
            out["R2AME_is"].append(trained["R2AME"])                                # :1300
            out["MSE_AME_is"].append(np.mean((y[tr] - X[tr] @ delta) ** 2))        # :1305
            yhat = X[te] @ delta
            out["R2AME_oos"].append(np.corrcoef(y[te], yhat)[0, 1] ** 2)           # :1311
            out["MSE_AME_oos"].append(np.mean((y[te] - yhat) ** 2))                # :1312
    return {k: np.array(v) for k, v in out.items()}


# --------------------------------------------------------------------------------------
# synthetic workload generator - shared by tests and bench
# --------------------------------------------------------------------------------------
def synthetic(N, P, seed, binary_last=False):
    rng = np.random.default_rng(seed)
    X0 = rng.standard_normal((N, P))
    eps = rng.standard_normal(N)
    y0 = np.sin(X0[:, 0]) + X0[:, 1] * X0[:, 2] + 0.5 * eps
    if binary_last:
        X0[:, P - 1] = (X0[:, P - 1] > 0.12345).astype(np.float64)
    return np.asfortranarray(X0), y0
