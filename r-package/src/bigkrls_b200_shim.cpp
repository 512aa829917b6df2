// .Call shim: the 11 native entry points of the reference package (src/RcppExports.cpp:147-160),
// re-implemented as thin forwards to libbigkrls_b200.so, plus _bigKRLS_fit / _bigKRLS_predict for the
// fused device-resident path.  Compiled only where R (>= 3.3) with Rcpp and bigmemory exists -
// NOT buildable in the development container (no R there); see INTEGRATION.md.
//
// Argument conventions are the reference's: matrices arrive as big.matrix@address external pointers
// (BigMatrix*: ->matrix(), ->nrow(), ->ncol(); column-major double, caller-allocated outputs written in
// place), vectors/scalars as plain R numerics; errors become R errors.
#include <Rcpp.h>
#include <bigmemory/BigMatrix.h>
#include "bigkrls_b200.h"

using namespace Rcpp;

static bk_ctx* ctx() {
  static bk_ctx* c = nullptr;
  if (!c) {
    const char* dev = std::getenv("BIGKRLS_DEVICE");
    if (bk_init(dev ? std::atoi(dev) : 0, &c) != BK_OK) stop(bk_last_error());
  }
  return c;
}
static void chk(int rc) {
  if (rc != BK_OK) stop(bk_last_error());
}
static double* mat(SEXP p) { return (double*)XPtr<BigMatrix>(p)->matrix(); }
static int64_t nr(SEXP p) { return XPtr<BigMatrix>(p)->nrow(); }
static int64_t nc(SEXP p) { return XPtr<BigMatrix>(p)->ncol(); }

// src/gauss_kernel.cpp:33-42
RcppExport SEXP _bigKRLS_BigGaussKernel(SEXP pA, SEXP pOut, SEXP sigma) {
  BEGIN_RCPP
  chk(bk_gauss_kernel(ctx(), mat(pA), nr(pA), nc(pA), as<double>(sigma), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
// src/temp_kernel.cpp:33-44
RcppExport SEXP _bigKRLS_BigTempKernel(SEXP pA, SEXP pB, SEXP pOut, SEXP sigma) {
  BEGIN_RCPP
  chk(bk_temp_kernel(ctx(), mat(pA), nr(pA), mat(pB), nr(pB), nc(pA), as<double>(sigma), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
// src/eigen.cpp:33-45  (Neig arrives as double; vals is 1 x Neig, vecs N x Neig, descending)
RcppExport SEXP _bigKRLS_BigEigen(SEXP pA, SEXP Neig, SEXP pVal, SEXP pVec) {
  BEGIN_RCPP
  const int64_t n = nr(pA), k = std::min<int64_t>(n, (int64_t)as<double>(Neig));
  chk(bk_eigen(ctx(), mat(pA), n, k, mat(pVal), mat(pVec)));
  return R_NilValue;
  END_RCPP
}
// src/solveforc.cpp:68-78  -> list(Le, coeffs)
RcppExport SEXP _bigKRLS_BigSolveForc(SEXP pQ, SEXP ev, SEXP y, SEXP lambda) {
  BEGIN_RCPP
  NumericVector e(ev), yy(y);
  const int64_t n = nr(pQ), k = nc(pQ);
  NumericMatrix coeffs(n, 1);  // arma::colvec wraps to an n x 1 matrix (src/solveforc.cpp:60-64)
  double Le = 0;
  chk(bk_solve_for_c(ctx(), mat(pQ), n, k, e.begin(), yy.begin(), as<double>(lambda), &Le, coeffs.begin()));
  return List::create(_["Le"] = Le, _["coeffs"] = coeffs);
  END_RCPP
}
// src/multdiag.cpp:26-37
RcppExport SEXP _bigKRLS_BigMultDiag(SEXP pA, SEXP diag, SEXP pOut) {
  BEGIN_RCPP
  NumericVector d(diag);
  chk(bk_mult_diag(ctx(), mat(pA), nr(pA), nc(pA), d.begin(), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
// src/crossprod.cpp:17-85
RcppExport SEXP _bigKRLS_BigCrossProd(SEXP pA, SEXP pB, SEXP pOut) {
  BEGIN_RCPP
  chk(bk_crossprod(ctx(), mat(pA), nr(pA), nc(pA), mat(pB), nc(pB), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
RcppExport SEXP _bigKRLS_BigXtX(SEXP pA, SEXP pOut) {
  BEGIN_RCPP
  chk(bk_xtx(ctx(), mat(pA), nr(pA), nc(pA), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
RcppExport SEXP _bigKRLS_BigTCrossProd(SEXP pA, SEXP pB, SEXP pOut) {
  BEGIN_RCPP
  chk(bk_tcrossprod(ctx(), mat(pA), nr(pA), nc(pA), mat(pB), nr(pB), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
RcppExport SEXP _bigKRLS_BigXXt(SEXP pA, SEXP pOut) {
  BEGIN_RCPP
  chk(bk_xxt(ctx(), mat(pA), nr(pA), nc(pA), mat(pOut)));
  return R_NilValue;
  END_RCPP
}
// src/bigderiv_v3.cpp:114-132
RcppExport SEXP _bigKRLS_BigDerivMat(SEXP pX, SEXP pK, SEXP pV, SEXP pD, SEXP pVar, SEXP coeffs, SEXP sigma) {
  BEGIN_RCPP
  NumericVector c(coeffs);
  chk(bk_deriv_mat(ctx(), mat(pX), nr(pX), nc(pX), mat(pK), mat(pV), c.begin(), as<double>(sigma), mat(pD),
                   mat(pVar)));
  return R_NilValue;
  END_RCPP
}
// src/Neffective.cpp:68-76
RcppExport SEXP _bigKRLS_BigNeffective(SEXP pX) {
  BEGIN_RCPP
  double out = 0;
  chk(bk_neffective(ctx(), mat(pX), nr(pX), nc(pX), &out));
  return wrap(out);
  END_RCPP
}

// ---- fused path: one .Call per fit, outputs written into caller-allocated big.matrix objects ---------
// The device-resident fit is owned by an external pointer: freed by R's garbage collector (or on any error
// path of the call that creates it), used again by _bigKRLS_predict.
struct FitGuard {
  bk_fit* f = nullptr;
  ~FitGuard() {
    if (f) bk_fit_free(f);
  }
  bk_fit* release() {
    bk_fit* r = f;
    f = nullptr;
    return r;
  }
};
static void fit_finalizer(SEXP xp) {
  bk_fit* f = (bk_fit*)R_ExternalPtrAddr(xp);
  if (f) bk_fit_free(f);
  R_ClearExternalPtr(xp);
}

// Xs, ys standardised (R/bigKRLS.R:251-254); pK / pVc / pVf / pD are big.matrix addresses or R_NilValue.
RcppExport SEXP _bigKRLS_fit(SEXP pXs, SEXP ys, SEXP sigma, SEXP Neig, SEXP eigtrunc, SEXP lambda, SEXP which,
                             SEXP derivative, SEXP vcov, SEXP ysd, SEXP pK, SEXP pVc, SEXP pVf, SEXP pD) {
  BEGIN_RCPP
  const int64_t n = nr(pXs), p = nc(pXs);
  NumericVector y(ys);
  IntegerVector wh(which);  // 1-based, length 0 = all
  std::vector<int32_t> w0(wh.size());
  for (int i = 0; i < wh.size(); ++i) w0[i] = wh[i] - 1;
  bk_fit_opts o;
  bk_fit_default_opts(&o, n, p);
  o.sigma = as<double>(sigma);
  o.neig = (int64_t)as<double>(Neig);
  o.eigtrunc = as<double>(eigtrunc);
  o.lambda = Rf_isNull(lambda) ? 0.0 : as<double>(lambda);
  o.derivative = as<bool>(derivative);
  o.vcov = as<bool>(vcov);
  o.n_which = (int)w0.size();
  o.which = w0.empty() ? nullptr : w0.data();
  o.y_sd = as<double>(ysd);
  if (!Rf_isNull(pK)) o.K_host = mat(pK);  // copied under the eigensolver; bk_fit_get_K below is then a no-op
  FitGuard g;  // chk() throws: the guard frees the device state on every error path below
  chk(bk_fit_run(ctx(), mat(pXs), y.begin(), n, p, &o, nullptr, &g.f));
  bk_fit_info info;
  chk(bk_fit_get_info(g.f, &info));
  NumericVector ev(info.neig), yhat(n), var(info.n_deriv);
  NumericMatrix coeffs(n, 1);
  chk(bk_fit_get_eigenvalues(g.f, ev.begin()));
  chk(bk_fit_get_coeffs(g.f, coeffs.begin()));
  chk(bk_fit_get_yfitted(g.f, yhat.begin()));
  if (!Rf_isNull(pK)) chk(bk_fit_get_K(g.f, mat(pK)));
  if (o.vcov && !Rf_isNull(pVc)) chk(bk_fit_get_vcov_c(g.f, mat(pVc)));
  if (o.vcov && !Rf_isNull(pVf)) chk(bk_fit_get_vcov_fitted(g.f, mat(pVf)));
  if (o.derivative && !Rf_isNull(pD)) {
    chk(bk_fit_get_derivatives(g.f, mat(pD)));
    chk(bk_fit_get_var_avgderiv(g.f, var.begin()));
  }
  SEXP handle = PROTECT(R_MakeExternalPtr(g.release(), R_NilValue, R_NilValue));
  R_RegisterCFinalizerEx(handle, fit_finalizer, TRUE);
  List out = List::create(_["values"] = ev, _["lastkeeper"] = (double)info.lastkeeper, _["lambda"] = info.lambda,
                          _["Le"] = info.Le, _["coeffs"] = coeffs, _["yfitted"] = yhat, _["varavgderiv"] = var,
                          _["handle"] = handle);
  UNPROTECT(1);
  return out;
  END_RCPP
}

// predict.bigKRLS on the device state of a fit (R/bigKRLS.R:596-613): newdata standardised with the training
// mean/sd; pKnew (M x N) and pVcovPred (M x M) are caller-allocated big.matrix addresses or R_NilValue.
// -> list(predicted (standardised units), se2 (diag of vcov.est.pred, before the Neffective correction))
RcppExport SEXP _bigKRLS_predict(SEXP handle, SEXP pNew, SEXP pKnew, SEXP pVcovPred, SEXP sePred) {
  BEGIN_RCPP
  bk_fit* f = (bk_fit*)R_ExternalPtrAddr(handle);
  if (!f) stop("bigKRLS fit handle was released; refit or use the per-op path");
  const int64_t m = nr(pNew);
  const bool se = as<bool>(sePred);
  NumericVector pred(m), se2(se ? m : 0);
  chk(bk_fit_predict_full(f, mat(pNew), m, pred.begin(), Rf_isNull(pKnew) ? nullptr : mat(pKnew),
                          se ? se2.begin() : nullptr, (se && !Rf_isNull(pVcovPred)) ? mat(pVcovPred) : nullptr));
  return List::create(_["predicted"] = pred, _["se2"] = se2);
  END_RCPP
}

static const R_CallMethodDef CallEntries[] = {
    {"_bigKRLS_BigNeffective", (DL_FUNC)&_bigKRLS_BigNeffective, 1},
    {"_bigKRLS_BigDerivMat", (DL_FUNC)&_bigKRLS_BigDerivMat, 7},
    {"_bigKRLS_BigCrossProd", (DL_FUNC)&_bigKRLS_BigCrossProd, 3},
    {"_bigKRLS_BigXtX", (DL_FUNC)&_bigKRLS_BigXtX, 2},
    {"_bigKRLS_BigTCrossProd", (DL_FUNC)&_bigKRLS_BigTCrossProd, 3},
    {"_bigKRLS_BigXXt", (DL_FUNC)&_bigKRLS_BigXXt, 2},
    {"_bigKRLS_BigEigen", (DL_FUNC)&_bigKRLS_BigEigen, 4},
    {"_bigKRLS_BigGaussKernel", (DL_FUNC)&_bigKRLS_BigGaussKernel, 3},
    {"_bigKRLS_BigMultDiag", (DL_FUNC)&_bigKRLS_BigMultDiag, 3},
    {"_bigKRLS_BigSolveForc", (DL_FUNC)&_bigKRLS_BigSolveForc, 4},
    {"_bigKRLS_BigTempKernel", (DL_FUNC)&_bigKRLS_BigTempKernel, 4},
    {"_bigKRLS_fit", (DL_FUNC)&_bigKRLS_fit, 14},
    {"_bigKRLS_predict", (DL_FUNC)&_bigKRLS_predict, 5},
    {NULL, NULL, 0}};

RcppExport void R_init_bigKRLS(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
