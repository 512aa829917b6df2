// CPU port of the reference's estimation path with the reference's OWN algorithmic structure
// (TEST INFRASTRUCTURE / CPU BASELINE ONLY - never linked into or called by bigkrls_b200/).
//
// This is the timed "reference CPU path" of bench.py (`cpu_baseline.kind = "port"`): the
// reference package itself cannot be built here (it needs R, Rcpp, RcppArmadillo, bigmemory -
// SURVEY.md section 8c), so its C++/R algorithm is restated literally over the same BLAS/LAPACK
// routines Armadillo dispatches to (OpenBLAS 0.3.x from the scipy wheel: dsyevd, dgemm, dgemv).
// The numpy twin is oracle/krls_oracle.py (literal=True); tests/test_oracle_port.py checks the two
// against each other and against the reference's golden vectors.
//
//   kernel       pairwise loop, j >= i, mirrored            src/gauss_kernel.cpp:18-23
//   eigen        dsyevd ("dc"), flip to descending          src/eigen.cpp:24-29
//                vecs <- -vecs, lastkeeper, truncation      R/bigKRLS_Rcpp_functions.R:186-196
//   solveforc    in-place transpose, N gemv's, transpose    src/solveforc.cpp:34-58
//   lambda       bounds + golden section, one solveforc per probe   R/bigKRLS_Rcpp_functions.R:10-77
//   fit          yfitted = K c, sigmasq, m = Q diag, m Q', K'(V K)  R/bigKRLS.R:286-307
//   derivatives  L = diff o K, L c, sum(L' V L) per column; binary branch
//                                                            src/bigderiv_v3.cpp:24-110
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <set>
#include <vector>

extern "C" {
void scipy_dsyevd_(const char* jobz, const char* uplo, const int* n, double* a, const int* lda,
                   double* w, double* work, const int* lwork, int* iwork, const int* liwork,
                   int* info);
void scipy_cblas_dgemm(int order, int ta, int tb, int m, int n, int k, double alpha, const double* A,
                       int lda, const double* B, int ldb, double beta, double* C, int ldc);
void scipy_cblas_dgemv(int order, int trans, int m, int n, double alpha, const double* A, int lda,
                       const double* x, int incx, double beta, double* y, int incy);
void scipy_openblas_set_num_threads(int n);
int scipy_openblas_get_num_threads(void);
}

namespace {
const int ColMajor = 102, NoTrans = 111, Trans = 112;
typedef std::vector<double> vec;
double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
void gemm(bool ta, bool tb, int m, int n, int k, double alpha, const double* A, int lda,
          const double* B, int ldb, double beta, double* C, int ldc) {
  scipy_cblas_dgemm(ColMajor, ta ? Trans : NoTrans, tb ? Trans : NoTrans, m, n, k, alpha, A, lda, B,
                    ldb, beta, C, ldc);
}

// src/solveforc.cpp:14-65 - literal structure (Q is n x k column-major on entry and exit)
void solveforc(vec& Q, int n, int k, const double* ev, const double* y, double lambda, double* Le,
               double* coeffs) {
  vec Qt((size_t)k * n);
  for (int j = 0; j < k; ++j)  // :34  Eigenvectors = trans(Eigenvectors)
    for (int i = 0; i < n; ++i) Qt[(size_t)i * k + j] = Q[(size_t)j * n + i];
  vec ginv_diag(n, 0.0), c(n, 0.0), row(k), ginv(n);
  for (int i = 0; i < n; ++i) {
    for (int j = 0; j < k; ++j) row[j] = Qt[(size_t)i * k + j] / (ev[j] + lambda);
    // ginv (1 x (i+1)) = row (1 x k) * temp_eigen (k x (i+1))      :42
    scipy_cblas_dgemv(ColMajor, Trans, k, i + 1, 1.0, Qt.data(), k, row.data(), 1, 0.0, ginv.data(), 1);
    ginv_diag[i] = ginv[i];                                          // :44
    double s = 0.0;
    for (int t = 0; t < i; ++t) c[t] += ginv[t] * y[i];              // :45
    for (int t = 0; t <= i; ++t) s += ginv[t] * y[t];                // :46
    c[i] += s;
  }
  for (int j = 0; j < k; ++j)  // :54  transpose back
    for (int i = 0; i < n; ++i) Q[(size_t)j * n + i] = Qt[(size_t)i * k + j];
  double le = 0.0;
  for (int i = 0; i < n; ++i) le += std::pow(c[i] / ginv_diag[i], 2);  // :56-58
  *Le = le;
  std::memcpy(coeffs, c.data(), sizeof(double) * n);
}
}  // namespace

extern "C" {

// times[8]: kernel, eigen, lambda, coeffs+fit, vcov, derivatives, total, #probes
// Xs (n x p), ys (n): STANDARDISED inputs (the R driver standardises before the native calls).
__attribute__((visibility("default"))) int krls_port_fit(
    const double* Xs, const double* ys, int n, int p, double sigma, double eigtrunc, int threads,
    int do_deriv, double* times, double* lambda_out, int* lastkeeper_out, double* evals_out,
    double* coeffs_out, double* yfitted_out, double* deriv_out, double* var_out, double* Le_out) {
  if (threads > 0) scipy_openblas_set_num_threads(threads);
  const double t_all = now();
  double t0 = now();
  // ---- 1/5 kernel (src/gauss_kernel.cpp:18-23) ----------------------------------------------
  vec K((size_t)n * n);
  {
    vec xi(p);
    for (int i = 0; i < n; ++i) {
      for (int d = 0; d < p; ++d) xi[d] = Xs[(size_t)d * n + i];
      for (int j = i; j < n; ++j) {
        double s = 0.0;
        for (int d = 0; d < p; ++d) {
          const double df = xi[d] - Xs[(size_t)d * n + j];
          s += df * df;
        }
        const double v = std::exp(-1 * s / sigma);
        K[(size_t)i * n + j] = v;
        K[(size_t)j * n + i] = v;
      }
    }
  }
  times[0] = now() - t0;
  // ---- 2/5 eigen (src/eigen.cpp:24-29; bEigen) ------------------------------------------------
  t0 = now();
  vec Q(K), w(n);
  {
    int lwork = -1, liwork = -1, info = 0, iwq = 0;
    double wq = 0;
    scipy_dsyevd_("V", "L", &n, Q.data(), &n, w.data(), &wq, &lwork, &iwq, &liwork, &info);
    lwork = (int)wq;
    liwork = iwq;
    vec work(lwork);
    std::vector<int> iwork(liwork);
    scipy_dsyevd_("V", "L", &n, Q.data(), &n, w.data(), work.data(), &lwork, iwork.data(), &liwork,
                  &info);
    if (info != 0) return info;
  }
  vec ev(n);
  for (int i = 0; i < n; ++i) ev[i] = w[n - 1 - i];                     // flipud
  int lastkeeper = 0;
  for (int i = 0; i < n; ++i)
    if (ev[i] >= eigtrunc * ev[0]) lastkeeper = i + 1;                  // R:190
  const int k = lastkeeper;
  vec Qk((size_t)n * k);
  for (int j = 0; j < k; ++j)                                            // fliplr, -1*vecs, deepcopy
    for (int i = 0; i < n; ++i) Qk[(size_t)j * n + i] = -Q[(size_t)(n - 1 - j) * n + i];
  vec().swap(Q);
  times[1] = now() - t0;
  // ---- 3/5 lambda (R/bigKRLS_Rcpp_functions.R:10-77) --------------------------------------------
  t0 = now();
  auto ratio = [&](double x) {
    long double s = 0;
    for (int i = 0; i < n; ++i) s += (long double)(ev[i] / (ev[i] + x));
    return (double)s;
  };
  double U = n;
  while (ratio(U) < 1) U -= 1;
  double L = 2.220446049250313e-16;
  int q = 0;
  {
    double best = INFINITY;
    for (int i = 0; i < n; ++i) {
      const double v = std::fabs(ev[i] - ev[0] / 1000);
      if (v < best) { best = v; q = i + 1; }
    }
  }
  while (ratio(L) > q) L += 0.05;
  vec ctmp(n);
  int probes = 0;
  auto loo = [&](double lam) {
    double le;
    solveforc(Qk, n, k, ev.data(), ys, lam, &le, ctmp.data());
    ++probes;
    return le;
  };
  const double tol = 1e-3 * n;
  double X1 = L + 0.381966 * (U - L), X2 = U - 0.381966 * (U - L);
  double S1 = loo(X1), S2 = loo(X2);
  while (std::fabs(S1 - S2) > tol) {
    if (S1 < S2) { U = X2; X2 = X1; X1 = L + 0.381966 * (U - L); S2 = S1; S1 = loo(X1); }
    else         { L = X1; X1 = X2; X2 = U - 0.381966 * (U - L); S1 = S2; S2 = loo(X2); }
  }
  const double lambda = (S1 < S2) ? X1 : X2;
  times[2] = now() - t0;
  times[7] = probes;
  // ---- 4/5 coefficients, fitted values (R/bigKRLS.R:286-294) ------------------------------------
  t0 = now();
  vec c(n), yfit(n);
  double Le;
  solveforc(Qk, n, k, ev.data(), ys, lambda, &Le, c.data());
  scipy_cblas_dgemv(ColMajor, NoTrans, n, n, 1.0, K.data(), n, c.data(), 1, 0.0, yfit.data(), 1);
  double sigmasq = 0;
  for (int i = 0; i < n; ++i) sigmasq += (ys[i] - yfit[i]) * (ys[i] - yfit[i]);
  sigmasq /= n;
  times[3] = now() - t0;
  // ---- vcov (R/bigKRLS.R:299-307) ----------------------------------------------------------------
  t0 = now();
  vec M((size_t)n * k), V((size_t)n * n), VK((size_t)n * n), Vyhat((size_t)n * n);
  for (int j = 0; j < k; ++j) {                                          // multdiag.cpp:17-18
    const double d = sigmasq * std::pow(ev[j] + lambda, -2);
    for (int i = 0; i < n; ++i) M[(size_t)j * n + i] = Qk[(size_t)j * n + i] * d;
  }
  gemm(false, true, n, n, k, 1.0, M.data(), n, Qk.data(), n, 0.0, V.data(), n);        // crossprod.cpp:53
  gemm(false, false, n, n, n, 1.0, V.data(), n, K.data(), n, 0.0, VK.data(), n);       // vcovmatc %*% K
  gemm(true, false, n, n, n, 1.0, K.data(), n, VK.data(), n, 0.0, Vyhat.data(), n);    // bCrossProd(K, .)
  vec().swap(VK);
  vec().swap(M);
  times[4] = now() - t0;
  // ---- 5/5 derivatives (src/bigderiv_v3.cpp:24-110) ------------------------------------------------
  t0 = now();
  if (do_deriv) {
    vec Lm((size_t)n * n), T1((size_t)n * n), T2((size_t)n * n);
    for (int j = 0; j < p; ++j) {
      const double* x = Xs + (size_t)j * n;
      std::set<double> uniq(x, x + n);
      if (uniq.size() == 2) {                                            // :31 binary
        const double z0 = *std::min_element(x, x + n), z1 = *std::max_element(x, x + n);
        const double sdXj = 1 / (z1 - z0), phi = -1 / (std::pow(sdXj, 2) * sigma);
        vec KT(n), KC(n), sT(n), sC(n);
        vec& MT = Lm;
        vec& MC = T1;
        for (int i = 0; i < n; ++i) {                                    // :49-78
          double kt = 0, kc = 0, dsum = 0;
          const int c1 = x[i] == z0;
          for (int t = 0; t < n; ++t) {
            const double bmax = (x[t] + x[i] == 2 * z1), bmin = (x[t] + x[i] == 2 * z0);
            const double fg = x[i] > x[t], sg = x[i] < x[t];
            const double aTl = bmin - fg, aCl = bmax - sg;
            MT[(size_t)t * n + i] = std::exp((aTl + fg - sg) * phi) * K[(size_t)t * n + i];  // row i of adj_T
            MC[(size_t)t * n + i] = std::exp((aCl - fg + sg) * phi) * K[(size_t)t * n + i];
            kt += std::exp(aTl * phi) * K[(size_t)i * n + t];
            kc += std::exp(aCl * phi) * K[(size_t)i * n + t];
            const double c2 = std::exp((-2 * (bmax + bmin) + 1) * std::pow(z1 - z0, 2) / sigma);
            dsum += sdXj * std::pow(-1, c1) * (1 - c2) * K[(size_t)i * n + t] * c[t];
          }
          KT[i] = kt;
          KC[i] = kc;
          deriv_out[(size_t)j * n + i] = dsum;
        }
        // :82-84  three N x N x N products, column sums
        gemm(false, true, n, n, n, 1.0, MT.data(), n, V.data(), n, 0.0, T2.data(), n);
        for (int t = 0; t < n; ++t) { double s = 0; for (int i = 0; i < n; ++i) s += T2[(size_t)t * n + i]; sT[t] = s; }
        gemm(false, true, n, n, n, 1.0, MC.data(), n, V.data(), n, 0.0, T2.data(), n);
        for (int t = 0; t < n; ++t) { double s = 0; for (int i = 0; i < n; ++i) s += T2[(size_t)t * n + i]; sC[t] = s; }
        gemm(false, true, n, n, n, 1.0, MT.data(), n, V.data(), n, 0.0, T2.data(), n);
        double vs = 0;
        for (int t = 0; t < n; ++t) {
          double s3 = 0;
          for (int i = 0; i < n; ++i) s3 += T2[(size_t)t * n + i];
          vs += sT[t] * KT[t] + sC[t] * KC[t] - 2 * s3 * KC[t];
        }
        var_out[j] = 2 * std::pow(sdXj, 2) * vs / std::pow((double)n, 2);  // :85
      } else {                                                           // :90-106 continuous
        for (int i = 0; i < n; ++i)
          for (int t = 0; t < n; ++t)
            Lm[(size_t)i * n + t] = (x[t] - x[i]) * K[(size_t)i * n + t];   // differences.col(i) % K
        scipy_cblas_dgemv(ColMajor, NoTrans, n, n, -2 / sigma, Lm.data(), n, c.data(), 1, 0.0,
                          deriv_out + (size_t)j * n, 1);
        gemm(true, false, n, n, n, 1.0, Lm.data(), n, V.data(), n, 0.0, T1.data(), n);   // L.t() * V
        gemm(false, false, n, n, n, 1.0, T1.data(), n, Lm.data(), n, 0.0, T2.data(), n); // (..) * L
        long double s = 0;
        for (size_t t = 0; t < (size_t)n * n; ++t) s += T2[t];
        var_out[j] = (1 / std::pow((double)n, 2)) * std::pow(-2 / sigma, 2) * (double)s;
      }
    }
  }
  times[5] = now() - t0;
  times[6] = now() - t_all;
  *lambda_out = lambda;
  *lastkeeper_out = lastkeeper;
  *Le_out = Le;
  std::memcpy(evals_out, ev.data(), sizeof(double) * n);
  std::memcpy(coeffs_out, c.data(), sizeof(double) * n);
  std::memcpy(yfitted_out, yfit.data(), sizeof(double) * n);
  return 0;
}

__attribute__((visibility("default"))) int krls_port_threads(void) { return scipy_openblas_get_num_threads(); }

}  // extern "C"
