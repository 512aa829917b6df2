"""Host side of the drop-in: the reference's user API, over the C ABI.

Mirrors R/bigKRLS.R of the reference (same argument names and meaning, same output-list field
names, same error messages where the reference validates):

    bigKRLS(y, X, sigma, derivative, which.derivatives, vcov.est, Neig, eigtrunc, lambda, L, U,
            tol, acf, ...)                                   R/bigKRLS.R:97-516
    predict(object, newdata, se.pred, correct_SE, ytest)     R/bigKRLS.R:547-637
    summary(object, degrees, probs, digits, labs)            R/bigKRLS.R:666-757
    crossvalidate_bigKRLS(y, X, seed, Kfolds, ptesting, ...) R/bigKRLS.R:1146-1336

Only the O(NP) glue the reference keeps in R (standardisation, rescaling by sd(y)/sd(X), R2,
R2AME - R/bigKRLS.R:245-254,390-407,420-453) lives here in numpy; everything O(N^2) and above
runs in libbigkrls_b200.so on the GPU.  R is not available in the build container, so this
Python module stands where the R wrappers would (see INTEGRATION.md for the `.Call` shim).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import BKError, FitInfo, FitOpts, check, dptr, fmat


def _col_sd(X):
    return np.std(X, axis=0, ddof=1)


class BigKRLS(dict):
    """Output object of bigKRLS(): a dict with the reference's field names
    (R/bigKRLS.R:148-151,268-269,280,392,420-469) plus the live device handle."""

    _fit = None
    _ctx = None

    def release_device(self):
        """Free the HBM held by this fit (K, Q, vcov blocks)."""
        if self._fit is not None:
            _lib.load().bk_fit_free(self._fit)
            self._fit = None

    def release_pinned(self):
        """Return pinned output buffers (pinned=True fits) to the context's pool."""
        for k in ("K", "vcov.est.c", "vcov.est.fitted"):
            a = self.pop(k, None)
            if a is not None and self._ctx is not None and self.get("_pinned"):
                self._ctx.recycle_pinned(a)

    def __del__(self):
        try:
            self.release_device()
        except Exception:
            pass


def _stop(msg):
    raise ValueError(msg)


def bigKRLS(y=None, X=None, sigma=None, derivative=True, which_derivatives=None, vcov_est=True,
            Neig=None, eigtrunc=None, lambda_=None, L=None, U=None, tol=None, acf=False,
            noisy=None, instructions=False, ctx=None, comm=None, return_squares=True,
            keep_device=True, loo_batch=15, fix_sd_index_bug=False, pinned=False, Ncores=None,
            model_subfolder_name=None, overwrite_existing=False, squares_alloc=None):
    """Kernel-regularised least squares, all heavy stages on the GPU.

    `which_derivatives` is 1-based like R.  `comm` is a bigkrls_b200.dist.TorchComm for
    multi-GPU (one process per GPU) runs.  The reference's quirk B.1 (derivative column i is
    divided by X.init.sd[i], not X.init.sd[which.derivatives[i]], R/bigKRLS.R:395-397) is
    replicated unless fix_sd_index_bug=True.  `Ncores` - in the reference the number of PSOCK worker processes of
    the marginal-effects stage (R/bigKRLS.R:257,332-363) - selects how many GPUs (ranks of `comm`) take part in the
    partitioned stages: the first Ncores ranks run the fit, the others return None; without `comm` there is one
    GPU and Ncores has nothing to select."""
    lib = _lib.load()
    sub_comm = None
    if comm is not None and Ncores is not None and int(Ncores) < comm.world:
        comm = sub_comm = comm.sub(int(Ncores))        # collective over all ranks of the parent communicator
        if comm is None:
            return None
    if X is None or y is None:
        _stop("y and X are required")
    X0 = np.asarray(X, dtype=np.float64)
    if X0.ndim != 2:
        _stop("is.matrix(X) | is.big.matrix(X) is not TRUE")                    # :146
    y0 = np.asarray(y, dtype=np.float64).reshape(-1)
    n, p = X0.shape
    w = BigKRLS()
    return_big = n > 2500                                                        # :150
    w["has.big.matrices"] = bool(return_big)
    xlabs = [f"x{i + 1}" for i in range(p)]                                     # :166-172
    w["X"] = X0
    X_init_sd = _col_sd(X0)                                                      # :179
    if np.isnan(X0).any():                                                       # :183-187
        bad = np.nonzero(np.isnan(X0).any(axis=0))[0] + 1
        _stop("the following columns in X contain missing data, which must be removed: "
              + ", ".join(str(int(b)) for b in bad))
    acf = bool(acf) and p > 2                                                    # :192
    Neig_ = min(n, int(Neig)) if Neig is not None else n                        # :194
    if eigtrunc is None:                                                         # :195-201
        eigtrunc = 0.001 if n > 3000 else 0.0
    elif not isinstance(eigtrunc, (int, float)) or eigtrunc < 0 or eigtrunc > 1:
        _stop("eigtrunc must be between 0 (no truncation) and 1 (keep largest only).")
    if which_derivatives is not None:                                            # :206-215
        if not derivative:
            _stop("which.derivative requires derivative = TRUE\n\nDerivative is a logical indicating "
                  "whether derivatives should be estimated (as opposed to just coefficients); "
                  "which.derivatives is a vector indicating which one (with NULL meaning all).")
        which_derivatives = [int(i) for i in which_derivatives]
        if not all(1 <= i <= p for i in which_derivatives):
            _stop("sum(which.derivatives %in% 1:p) == length(which.derivatives) is not TRUE")
    if X_init_sd.min() == 0:                                                     # :217-218
        _stop("The following columns in X are constant and must be removed: "
              + " ".join(str(int(i) + 1) for i in np.nonzero(X_init_sd == 0)[0]))
    if n != y0.shape[0]:
        _stop("nrow(X) not equal to number of elements in y.")                  # :219-220
    if np.isnan(y0).any():
        _stop("y contains missing data.")                                       # :221-222
    y_sd = float(np.std(y0, ddof=1))
    if y_sd == 0:
        _stop("y is a constant.")                                               # :223-224
    if lambda_ is not None and not (lambda_ > 0):
        _stop("lambda > 0 is not TRUE")                                         # :225-226
    if sigma is not None and not (sigma > 0):
        _stop("sigma > 0 is not TRUE")                                          # :227-228
    sigma = float(p) if sigma is None else float(sigma)                         # :230
    if derivative and not vcov_est:                                              # :238-239
        _stop("vcov.est is needed to get derivatives (derivative==TRUE requires vcov.est=TRUE).")
    x_is_binary = np.array([np.unique(X0[:, j]).size == 2 for j in range(p)])   # :242
    y_mean = float(y0.mean())
    Xs = fmat((X0 - X0.mean(axis=0)) / X_init_sd)                               # :251-253
    ys = np.ascontiguousarray((y0 - y_mean) / y_sd)                             # :254

    ctx = ctx or _lib.default_context()
    opts = FitOpts()
    lib.bk_fit_default_opts(C.byref(opts), n, p)
    opts.sigma = sigma
    opts.eigtrunc = float(eigtrunc)
    opts.neig = Neig_
    opts.lambda_ = float(lambda_) if lambda_ is not None else 0.0
    opts.L = float(L) if L is not None else -1.0
    opts.U = float(U) if U is not None else 0.0
    opts.tol = 0.0          # the reference computes tol (:232-236) but never forwards it (:274-275)
    opts.derivative = 1 if derivative else 0
    opts.vcov = 1 if vcov_est else 0
    opts.y_sd = y_sd
    opts.loo_batch = int(loo_batch)
    opts.keep_vcov_fitted = 1 if vcov_est else 0
    wd0 = None
    if which_derivatives is not None:
        wd0 = np.ascontiguousarray(np.array(which_derivatives, dtype=np.int32) - 1)
        opts.n_which = len(which_derivatives)
        opts.which = wd0.ctypes.data_as(_lib.c_int32_p)
    h = C.c_void_p()
    cptr = C.byref(comm.struct) if comm is not None and comm.world > 1 else None
    # where the N x N outputs go: library-pinned buffers, plain (pageable) numpy memory, or anything the caller
    # provides through squares_alloc(name, shape) -> F-ordered float64 array - e.g. np.memmap files standing in for
    # the reference's file-backed big.matrix (R/bigKRLS_Rcpp_functions.R:143-147); pageable targets are filled by
    # the library's bounce-buffer copy engine (csrc/hostcopy.cu)
    if squares_alloc is not None:
        alloc = squares_alloc
    elif pinned:
        alloc = lambda name, shape: ctx.pinned_empty(shape)
    else:
        alloc = lambda name, shape: np.empty(shape, order="F")
    K = None
    if return_squares:
        # the kernel matrix is final after the first stage: hand its host buffer to the library so that the
        # device->host copy runs under the eigensolver instead of after the fit
        rank_, world_ = (comm.rank, comm.world) if cptr is not None else (0, 1)
        K = alloc("K", (n, n * (rank_ + 1) // world_ - n * rank_ // world_))
        opts.K_host = K.ctypes.data
    check(lib.bk_fit_run(ctx.handle, dptr(Xs), dptr(ys), n, p, C.byref(opts), cptr, C.byref(h)))
    w._fit, w._ctx = h, ctx
    if sub_comm is not None and getattr(sub_comm, "close", None):
        sub_comm.close()                               # the fit's device state does not live in the peer heaps
    info = FitInfo()
    check(lib.bk_fit_get_info(h, C.byref(info)))
    w["_info"] = info.as_dict()

    ev = np.empty(Neig_)
    check(lib.bk_fit_get_eigenvalues(h, dptr(ev)))
    w["K.eigenvalues"] = ev                                                     # :268
    w["lastkeeper"] = int(info.lastkeeper)                                      # :269
    lam = float(info.lambda_)
    w["Neffective"] = float(n - np.sum(ev / (ev + lam)))                        # :280
    coeffs = np.empty(n)
    check(lib.bk_fit_get_coeffs(h, dptr(coeffs)))
    yfit_std = np.empty(n)
    check(lib.bk_fit_get_yfitted(h, dptr(yfit_std)))

    c0, c1 = C.c_int64(), C.c_int64()
    check(lib.bk_fit_col_range(h, C.byref(c0), C.byref(c1)))
    w["_col_range"] = (c0.value, c1.value)
    ncols = c1.value - c0.value

    if derivative:                                                               # :321-407
        pd = int(info.n_deriv)
        D = np.empty((n, pd), order="F")
        check(lib.bk_fit_get_derivatives(h, dptr(D)))
        varavg = np.empty(pd)
        check(lib.bk_fit_get_var_avgderiv(h, dptr(varavg)))
        wsel = list(range(p)) if which_derivatives is None else [i - 1 for i in which_derivatives]
        X_est = Xs[:, wsel]
        yhat_ame = X_est @ D.mean(axis=0)                                       # :388
        w["R2AME"] = float(np.corrcoef(y0, yhat_ame)[0, 1] ** 2)                # :390
        D = y_sd * D                                                            # :392
        for i in range(pd):                                                     # :393-395
            D[:, i] = D[:, i] / (X_init_sd[wsel[i]] if fix_sd_index_bug else X_init_sd[i])
        w["avgderivatives"] = D.mean(axis=0)[None, :]                           # :398
        w["var.avgderivatives"] = ((y_sd / X_init_sd[wsel]) ** 2 * varavg)[None, :]   # :401-405
        w["derivatives"] = D
    if acf:                                                                      # :412-416
        out = C.c_double()
        check(lib.bk_neffective(ctx.handle, dptr(Xs), n, p, C.byref(out)))
        w["Neffective.acf"] = out.value

    w["coeffs"] = coeffs.reshape(-1, 1)                                         # :420
    w["y"] = y0
    w["sigma"] = sigma
    w["lambda"] = lam
    w["binaryindicator"] = x_is_binary
    w["which.derivatives"] = which_derivatives
    w["xlabs"] = xlabs
    w["yfitted"] = yfit_std * y_sd + y_mean                                     # :429
    w["R2"] = float(1 - np.var(y0 - w["yfitted"], ddof=1) / y_sd ** 2)          # :430
    w["Looe"] = float(info.Le) * y_sd                                           # :431
    if return_squares:
        assert K.shape == (n, ncols)
        check(lib.bk_fit_get_K(h, dptr(K)))                                     # no-op: delivered during the fit
        w["K"] = K                                                              # :435
        if vcov_est:
            Vc = alloc("vcov.est.c", (n, ncols))
            check(lib.bk_fit_get_vcov_c(h, dptr(Vc)))
            w["vcov.est.c"] = Vc                                                # :439-452
            Vf = alloc("vcov.est.fitted", (n, ncols))
            check(lib.bk_fit_get_vcov_fitted(h, dptr(Vf)))
            w["vcov.est.fitted"] = Vf
    w["_pinned"] = bool(pinned)
    w["derivative.call"] = bool(derivative)                                     # :455
    if not keep_device:
        w.release_device()
    return w


def _predict_rows(object, news, se_pred, ctx):
    """Rows `news` (standardised) -> (pred_std, Knew, se2 | None, G | None) on this process's GPU.

    With the fit's device state alive: the fused path (bk_fit_predict_full, spectral quadratic forms).  Without it
    (keep_device=False, released, or a fit rebuilt from saved fields) the reference's own recipe on the host
    fields X, coeffs, vcov.est.c through the per-op entry points (R/bigKRLS.R:599-608)."""
    lib = _lib.load()
    X = np.asarray(object["X"], dtype=np.float64)
    m, n = news.shape[0], X.shape[0]
    pred = np.empty(m)
    Knew = np.empty((m, n), order="F")
    se2 = np.empty(m) if se_pred else None
    vp = np.empty((m, m), order="F") if se_pred else None
    if object._fit is not None:
        check(lib.bk_fit_predict_full(object._fit, dptr(news), m, dptr(pred), dptr(Knew),
                                      dptr(se2) if se_pred else None, dptr(vp) if se_pred else None))
        return pred, Knew, se2, vp
    h = ctx.handle
    Xs = fmat((X - X.mean(axis=0)) / _col_sd(X))
    check(lib.bk_temp_kernel(h, dptr(news), m, dptr(Xs), n, X.shape[1], float(object["sigma"]), dptr(Knew)))
    c = np.ascontiguousarray(object["coeffs"], dtype=np.float64).reshape(-1)
    check(lib.bk_dgemm(h, 0, 0, m, 1, n, dptr(Knew), m, dptr(c), n, dptr(pred), m))
    if se_pred:
        V = object["vcov.est.c"]
        if V.shape != (n, n):
            raise BKError(-4, "predict(se.pred=TRUE) from host fields needs the full vcov.est.c (this is a column "
                              "block of a multi-GPU fit)")
        V = np.asfortranarray(V)
        T = np.empty((m, n), order="F")
        check(lib.bk_dgemm(h, 0, 0, m, n, n, dptr(Knew), m, dptr(V), n, dptr(T), m))      # newdataK %*% vcov.est.c
        check(lib.bk_dgemm(h, 0, 1, m, m, n, dptr(T), m, dptr(Knew), m, dptr(vp), m))     # bTCrossProd(., newdataK)
        se2 = np.diag(vp).copy()
    return pred, Knew, se2, vp


def predict(object, newdata, se_pred=False, correct_SE=True, ytest=None, comm=None):
    """predict.bigKRLS (R/bigKRLS.R:547-637).  Returns the reference's list fields: predicted, se.pred,
    vcov.est.pred, newdata, newdataK, has.big.matrices, ytest.

    With a multi-rank `comm` the rows of newdata are sharded over the ranks (independent rows);
    every rank returns the gathered full result."""
    if not isinstance(object, BigKRLS):
        raise TypeError("Object not of class 'bigKRLS'")
    X = np.asarray(object["X"], dtype=np.float64)
    new = np.asarray(newdata, dtype=np.float64)
    if new.ndim == 1:
        new = new.reshape(1, -1)
    if X.shape[1] != new.shape[1]:
        _stop("ncol(newdata) differs from ncol(X) from fitted bigKRLS object")  # :583-584
    if se_pred and object.get("vcov.est.c") is None and object._fit is None:
        _stop("recompute bigKRLS object with bigKRLS(,vcov.est=TRUE) to compute standard errors")
    Xmeans, Xsd = X.mean(axis=0), _col_sd(X)                                    # :587-588
    news = fmat((new - Xmeans) / Xsd)                                           # :593-594
    m, n = news.shape[0], X.shape[0]
    y = np.asarray(object["y"]).reshape(-1)
    ctx = object._ctx or _lib.default_context()
    world, rank = (comm.world, comm.rank) if comm is not None else (1, 0)
    if world > 1:
        r0, r1 = m * rank // world, m * (rank + 1) // world
        part = _predict_rows(object, fmat(news[r0:r1]), se_pred, ctx) if r1 > r0 else None
        parts = comm.gather_objects((r0, r1, part))
        pred, Knew = np.empty(m), np.empty((m, n), order="F")
        se2 = np.empty(m) if se_pred else None
        for a, b, pt in parts:
            if pt is not None:
                pred[a:b], Knew[a:b] = pt[0], pt[1]
                if se_pred:
                    se2[a:b] = pt[2]
        vp = None
        if se_pred:
            # off-diagonal blocks couple rows of different ranks: the M x M matrix from the gathered kernel rows
            _, _, _, vp = _predict_rows(object, news, True, ctx) if rank == 0 else (None, None, None, None)
            vp = comm.broadcast_object(vp, 0)
    else:
        pred, Knew, se2, vp = _predict_rows(object, news, se_pred, ctx)
    out = {"predicted": pred * np.std(y, ddof=1) + y.mean(),                    # :618
           "se.pred": None, "vcov.est.pred": None, "newdata": newdata, "newdataK": Knew,
           "has.big.matrices": object["has.big.matrices"], "ytest": ytest}
    if se_pred:
        v = se2.copy()
        if correct_SE and object.get("Neffective") is not None:
            f = math.sqrt(n / object["Neffective"])                             # :610-611
            v, vp = f * v, f * vp
        out["vcov.est.pred"] = vp                                               # :605
        out["se.pred"] = np.sqrt(v).reshape(-1, 1)                              # :613
    return out


def summary(object, degrees="Neffective", probs=(0.05, 0.25, 0.5, 0.75, 0.95), digits=4, labs=None,
            quiet=True):
    """summary.bigKRLS (R/bigKRLS.R:666-757): t-tests of the average marginal effects and
    quantiles of the pointwise ones."""
    from scipy import stats
    if not isinstance(object, BigKRLS):
        raise TypeError("Object not of class 'bigKRLS'")
    if degrees not in ("acf", "Neffective", "N"):
        _stop('degrees %in% c("acf", "Neffective", "N") is not TRUE')
    X = np.asarray(object["X"])
    N = n = X.shape[0]
    p = X.shape[1]
    if degrees == "Neffective":
        n = object["Neffective"]
    if degrees == "acf":
        if object.get("Neffective.acf") is None:
            lib = _lib.load()
            Xs = fmat((X - X.mean(axis=0)) / _col_sd(X))
            out = C.c_double()
            check(lib.bk_neffective((object._ctx or _lib.default_context()).handle, dptr(Xs), N, p,
                                    C.byref(out)))
            n = out.value                                                       # :682-690
        else:
            n = object["Neffective.acf"]
    if object.get("derivatives") is None:
        return None
    wd = object["which.derivatives"] or list(range(1, p + 1))
    est = object["avgderivatives"].reshape(-1)
    se = np.sqrt(object["var.avgderivatives"].reshape(-1))
    if degrees != "Neffective":
        se = se * N / n                                                         # :728-730
    tval = est / se
    pval = 2 * stats.t.sf(np.abs(tval), n - p)                                  # :732
    names = list(labs) if labs is not None else list(object["xlabs"])
    rows = []
    for i, j in enumerate(wd):
        nm = names[j - 1] + ("*" if object["binaryindicator"][j - 1] else "")
        rows.append(nm)
    AME = np.column_stack([est, se, tval, pval])
    qd = np.quantile(object["derivatives"], probs, axis=0).T                    # :746
    ans = {"ttests": AME, "percentiles": qd, "rownames": rows,
           "colnames": ["Estimate", "Std. Error", "t value", "Pr(>|t|)"],
           "lambda": object["lambda"], "N": N, "Neffective": n, "R2": object["R2"],
           "R2AME": object.get("R2AME")}
    if not quiet:
        print("\n\nMODEL SUMMARY:\n")
        print("lambda:", round(object["lambda"], digits))
        print("N:", N)
        if n != N:
            print("N Effective:", n)
        print("R2:", round(object["R2"], digits))
        if object.get("R2AME") is not None:
            print("R2AME**:", round(object["R2AME"], digits), "\n")
        print("Average Marginal Effects:\n")
        for r, row in zip(rows, np.round(AME, digits)):
            print(f"{r:>12s}", *row)
        print("\n\nPercentiles of Marginal Effects:\n")
        for r, row in zip(rows, np.round(qd, digits)):
            print(f"{r:>12s}", *row)
    return ans


def _fold_stats(y, X, tr, te, comm=None, **kw):
    """One fold: bigKRLS on the training rows + predict on the test rows + the fit statistics
    of R/bigKRLS.R:1293-1312."""
    kw = dict(kw)
    kw.setdefault("return_squares", False)
    trained = bigKRLS(y[tr], X[tr], comm=comm, **kw)
    tested = predict(trained, X[te])
    st = {"R2_is": trained["R2"],
          "R2_oos": float(np.corrcoef(y[te], tested["predicted"])[0, 1] ** 2),
          "MSE_is": float(np.mean((y[tr] - trained["yfitted"]) ** 2)),
          "MSE_oos": float(np.mean((y[te] - tested["predicted"]) ** 2))}
    if trained.get("avgderivatives") is not None:
        delta = trained["avgderivatives"].reshape(-1)
        cols = list(range(X.shape[1])) if trained["which.derivatives"] is None else \
            [i - 1 for i in trained["which.derivatives"]]
        st["R2AME_is"] = trained["R2AME"]
        st["MSE_AME_is"] = float(np.mean((y[tr] - X[tr][:, cols] @ delta) ** 2))
        yhat = X[te][:, cols] @ delta
        st["R2AME_oos"] = float(np.corrcoef(y[te], yhat)[0, 1] ** 2)
        st["MSE_AME_oos"] = float(np.mean((y[te] - yhat) ** 2))
    trained.release_device()
    return st, trained, tested


def crossvalidate_bigKRLS(y, X, seed=None, Kfolds=None, ptesting=None, folds=None, keep_models=False,
                          comm=None, **kw):
    """crossvalidate.bigKRLS (R/bigKRLS.R:1146-1336).

    Fold assignment: the reference draws `as.integer(cut(sample(N), breaks = Kfolds))` with
    R's RNG (:1232); pass that vector as `folds` for exact parity.  Without it numpy's
    Generator(seed) permutation is used.  With a multi-rank `comm`, folds are independent units
    sharded round-robin over ranks (no data-path collective) and the statistics are gathered."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).reshape(-1)
    N = X.shape[0]
    if (Kfolds is None) + (ptesting is None) + (folds is None) != 2:
        _stop("Specify either Kfolds or ptesting but not both.")                # :1148
    rng = np.random.default_rng(seed)
    if ptesting is not None:                                                     # :1173-1226
        if ptesting < 0 or ptesting > 100:
            _stop("ptesting, the percentage of data to be used for validation, must be between 0 and 100.")
        Ntesting = int(round(N * ptesting / 100))
        train = np.sort(rng.permutation(N)[: N - Ntesting])
        tr = np.zeros(N, bool)
        tr[train] = True
        st, trained, tested = _fold_stats(y, X, tr, ~tr, **kw)
        out = {"type": "crossvalidated", "seed": seed, "ptesting": ptesting,
               "indices": {"train.set": np.nonzero(tr)[0] + 1, "test.set": np.nonzero(~tr)[0] + 1},
               "pseudoR2_is": st["R2_is"], "pseudoR2_oos": st["R2_oos"], "MSE_is": st["MSE_is"],
               "MSE_oos": st["MSE_oos"], "trained": trained, "tested": tested}
        for a, b in (("pseudoR2AME_is", "R2AME_is"), ("pseudoR2AME_oos", "R2AME_oos"),
                     ("MSE_AME_is", "MSE_AME_is"), ("MSE_AME_oos", "MSE_AME_oos")):
            if b in st:
                out[a] = st[b]
        return out
    if folds is None:
        if not (Kfolds > 0 and Kfolds % 1 == 0):
            _stop("is.numeric(Kfolds) & Kfolds > 0 & Kfolds%%1 == 0 is not TRUE")
        perm = rng.permutation(N)
        folds = np.empty(N, dtype=np.int64)
        # cut(sample(N), breaks = K): equal-width bins over the permuted ranks
        edges = np.linspace(1 - (N - 1) * 0.001, N + (N - 1) * 0.001, int(Kfolds) + 1)
        folds[:] = np.clip(np.searchsorted(edges, perm + 1, side="left"), 1, int(Kfolds))
    folds = np.asarray(folds).astype(np.int64)
    Kf = int(folds.max())
    for k in range(1, Kf + 1):                                                   # :1234-1243 check_data
        Xtr = X[folds != k]
        if np.std(Xtr, axis=0, ddof=1).min() == 0:
            _stop("The following columns in X are constant and must be removed: "
                  + " ".join(str(int(i) + 1) for i in np.nonzero(np.std(Xtr, axis=0, ddof=1) == 0)[0]))
    rank, world = (comm.rank, comm.world) if comm is not None else (0, 1)
    keys = ["R2_is", "R2_oos", "MSE_is", "MSE_oos", "R2AME_is", "R2AME_oos", "MSE_AME_is", "MSE_AME_oos"]
    local = {}
    models = {}
    for k in range(1, Kf + 1):
        if (k - 1) % world != rank:
            continue
        st, trained, tested = _fold_stats(y, X, folds != k, folds == k, **kw)
        local[k] = st
        if keep_models:
            models[k] = (trained, tested)
    allst = comm.gather_objects(local) if world > 1 else [local]
    merged = {}
    for d in allst:
        merged.update(d)
    out = {"type": "KfoldsCV", "Kfolds": Kf, "seed": seed, "folds": folds}
    for key in keys:
        vals = [merged[k].get(key) for k in range(1, Kf + 1)]
        if all(v is not None for v in vals):
            out[key] = np.array(vals)
    for k, (trained, tested) in models.items():
        out[f"fold_{k}"] = {"trained": trained, "tested": tested}
    return out
