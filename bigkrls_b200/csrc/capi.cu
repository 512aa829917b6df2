// Per-op C ABI on HOST buffers: one entry point per native export of the reference
// (R/RcppExports.R:4-46; src/RcppExports.cpp:147-160).  Each call stages its operands in HBM,
// runs the CUDA path and copies the result back.  There is no CPU fallback.
#include <cmath>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"

using namespace bk;

namespace {

int h2d(bk_ctx* ctx, DevBuf<double>& buf, const double* host, size_t n) {
  BK_TRY(buf.alloc(n));
  BK_CUDA(cudaMemcpyAsync(buf.p, host, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->stream));
  return BK_OK;
}
int d2h(bk_ctx* ctx, double* host, const double* dev, size_t n) {
  BK_CUDA(cudaMemcpyAsync(host, dev, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}
int enter(bk_ctx* ctx, const char* fn) {
  if (!ctx) {
    set_error("%s: ctx is NULL", fn);
    return BK_ERR_ARG;
  }
  BK_CUDA(bk::bind_ctx(ctx));
  return BK_OK;
}
bool fits_int(int64_t v) { return v >= 0 && v < 2147483647LL; }

}  // namespace

extern "C" {

int bk_gauss_kernel(bk_ctx* ctx, const double* X, int64_t n, int64_t p, double sigma, double* K) {
  BK_TRY(enter(ctx, "bk_gauss_kernel"));
  BK_REQUIRE(X && K && n > 0 && p > 0 && fits_int(n) && fits_int(p), "bk_gauss_kernel: bad arguments");
  BK_REQUIRE(sigma > 0.0, "bk_gauss_kernel: sigma must be positive");
  DevBuf<double> dX, dK;
  BK_TRY(h2d(ctx, dX, X, (size_t)n * p));
  BK_TRY(dK.alloc((size_t)n * n));
  BK_TRY(gauss_kernel_sym(ctx, dX.p, n, (int)n, (int)p, sigma, dK.p, n));
  return d2h(ctx, K, dK.p, (size_t)n * n);
}

int bk_temp_kernel(bk_ctx* ctx, const double* A, int64_t m, const double* B, int64_t n, int64_t p,
                   double sigma, double* out) {
  BK_TRY(enter(ctx, "bk_temp_kernel"));
  BK_REQUIRE(A && B && out && m > 0 && n > 0 && p > 0 && fits_int(m) && fits_int(n),
             "bk_temp_kernel: bad arguments");
  BK_REQUIRE(sigma > 0.0, "bk_temp_kernel: sigma must be positive");
  DevBuf<double> dA, dB, dO;
  BK_TRY(h2d(ctx, dA, A, (size_t)m * p));
  BK_TRY(h2d(ctx, dB, B, (size_t)n * p));
  BK_TRY(dO.alloc((size_t)m * n));
  BK_TRY(gauss_kernel_rect(ctx, dA.p, m, (int)m, dB.p, n, (int)n, (int)p, sigma, dO.p, m));
  return d2h(ctx, out, dO.p, (size_t)m * n);
}

int bk_eigen(bk_ctx* ctx, const double* A, int64_t n, int64_t neig, double* vals, double* vecs) {
  BK_TRY(enter(ctx, "bk_eigen"));
  BK_REQUIRE(A && vals && n > 0 && fits_int(n), "bk_eigen: bad arguments");
  BK_REQUIRE(neig >= 1 && neig <= n, "bk_eigen: neig must be in 1..n");
  DevBuf<double> dA, Z;
  BK_TRY(h2d(ctx, dA, A, (size_t)n * n));
  if (vecs) BK_TRY(Z.alloc((size_t)n * neig));
  if (use_topk(n, neig)) {
    // Neig << N: block-Krylov path (reference: sp_mat + eigs_sym, src/eigen.cpp:18-22)
    std::vector<double> evk(neig);
    BK_TRY(eigen_topk(ctx, dA.p, n, (int)n, (int)neig, evk.data(), vecs ? Z.p : nullptr, n, nullptr));
    for (int64_t i = 0; i < neig; ++i) vals[i] = evk[i];
    if (vecs) return d2h(ctx, vecs, Z.p, (size_t)n * neig);
    return BK_OK;
  }
  std::vector<double> ev(n);
  int nw = 0;
  // rel_thresh = -inf: keep all neig leading vectors (truncation is the caller's business,
  // R/bigKRLS_Rcpp_functions.R:190)
  BK_TRY(eigen_full(ctx, dA.p, n, (int)n, ev.data(), vecs ? (int)neig : 0, -INFINITY, &nw,
                    vecs ? Z.p : nullptr, n, nullptr));
  for (int64_t i = 0; i < neig; ++i) vals[i] = ev[i];
  if (vecs) return d2h(ctx, vecs, Z.p, (size_t)n * neig);
  return BK_OK;
}

int bk_loo_batch(bk_ctx* ctx, const double* Q, int64_t n, int64_t k, const double* evals,
                 const double* y, const double* lambdas, int nlam, double* Le) {
  BK_TRY(enter(ctx, "bk_loo_batch"));
  BK_REQUIRE(Q && evals && y && lambdas && Le && n > 0 && k > 0 && fits_int(n) && fits_int(k),
             "bk_loo_batch: bad arguments");
  BK_REQUIRE(nlam >= 1 && nlam <= 16, "bk_loo_batch: nlam must be in 1..16");
  DevBuf<double> dQ, dev, dy, dz, dLe;
  BK_TRY(h2d(ctx, dQ, Q, (size_t)n * k));
  BK_TRY(h2d(ctx, dev, evals, (size_t)k));
  BK_TRY(h2d(ctx, dy, y, (size_t)n));
  BK_TRY(dz.alloc(k));
  BK_TRY(dLe.alloc(16));
  BK_TRY(gemm(ctx, true, false, (int)k, 1, (int)n, 1.0, dQ.p, n, dy.p, n, 0.0, dz.p, k));
  BK_TRY(loo_batch(ctx, dQ.p, n, (int)n, (int)k, dev.p, dz.p, lambdas, nlam, dLe.p, nullptr));
  return d2h(ctx, Le, dLe.p, nlam);
}

int bk_solve_for_c(bk_ctx* ctx, const double* Q, int64_t n, int64_t k, const double* evals,
                   const double* y, double lambda, double* Le, double* coeffs) {
  BK_TRY(enter(ctx, "bk_solve_for_c"));
  BK_REQUIRE(Q && evals && y && Le && coeffs && n > 0 && k > 0 && fits_int(n) && fits_int(k),
             "bk_solve_for_c: bad arguments");
  DevBuf<double> dQ, dev, dy, dz, dLe, dc;
  BK_TRY(h2d(ctx, dQ, Q, (size_t)n * k));
  BK_TRY(h2d(ctx, dev, evals, (size_t)k));
  BK_TRY(h2d(ctx, dy, y, (size_t)n));
  BK_TRY(dz.alloc(k));
  BK_TRY(dLe.alloc(16));
  BK_TRY(dc.alloc(n));
  BK_TRY(gemm(ctx, true, false, (int)k, 1, (int)n, 1.0, dQ.p, n, dy.p, n, 0.0, dz.p, k));
  BK_TRY(loo_batch(ctx, dQ.p, n, (int)n, (int)k, dev.p, dz.p, &lambda, 1, dLe.p, dc.p));
  BK_TRY(d2h(ctx, Le, dLe.p, 1));
  return d2h(ctx, coeffs, dc.p, n);
}

int bk_mult_diag(bk_ctx* ctx, const double* A, int64_t n, int64_t k, const double* diag,
                 double* out) {
  BK_TRY(enter(ctx, "bk_mult_diag"));
  BK_REQUIRE(A && diag && out && n > 0 && k > 0 && fits_int(n) && fits_int(k),
             "bk_mult_diag: bad arguments");
  DevBuf<double> dA, dd, dO;
  BK_TRY(h2d(ctx, dA, A, (size_t)n * k));
  BK_TRY(h2d(ctx, dd, diag, (size_t)k));
  BK_TRY(dO.alloc((size_t)n * k));
  BK_TRY(col_scale(ctx, dA.p, n, (int)n, (int)k, dd.p, nullptr, dO.p, n));
  return d2h(ctx, out, dO.p, (size_t)n * k);
}

int bk_dgemm(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, const double* A,
             int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc) {
  BK_TRY(enter(ctx, "bk_dgemm"));
  BK_REQUIRE(A && B && C && m > 0 && n > 0 && k > 0 && fits_int(m) && fits_int(n) && fits_int(k),
             "bk_dgemm: bad arguments");
  const int64_t a_cols = ta ? m : k, b_cols = tb ? k : n;
  BK_REQUIRE(lda >= (ta ? k : m) && ldb >= (tb ? n : k) && ldc >= m, "bk_dgemm: bad leading dimension");
  DevBuf<double> dA, dB, dC;
  BK_TRY(h2d(ctx, dA, A, (size_t)lda * a_cols));
  BK_TRY(h2d(ctx, dB, B, (size_t)ldb * b_cols));
  BK_TRY(dC.alloc((size_t)ldc * n));
  BK_CUDA(cudaMemsetAsync(dC.p, 0, sizeof(double) * (size_t)ldc * n, ctx->stream));
  BK_TRY(gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, 1.0, dA.p, lda, dB.p, ldb, 0.0, dC.p, ldc));
  return d2h(ctx, C, dC.p, (size_t)ldc * n);
}

int bk_crossprod(bk_ctx* ctx, const double* A, int64_t r, int64_t m, const double* B, int64_t n,
                 double* out) {
  return bk_dgemm(ctx, 1, 0, m, n, r, A, r, B, r, out, m);
}
int bk_xtx(bk_ctx* ctx, const double* A, int64_t r, int64_t m, double* out) {
  return bk_dgemm(ctx, 1, 0, m, m, r, A, r, A, r, out, m);
}
int bk_tcrossprod(bk_ctx* ctx, const double* A, int64_t m, int64_t r, const double* B, int64_t n,
                  double* out) {
  return bk_dgemm(ctx, 0, 1, m, n, r, A, m, B, n, out, m);
}
int bk_xxt(bk_ctx* ctx, const double* A, int64_t m, int64_t r, double* out) {
  return bk_dgemm(ctx, 0, 1, m, m, r, A, m, A, m, out, m);
}

int bk_deriv_mat(bk_ctx* ctx, const double* X, int64_t n, int64_t p, const double* K,
                 const double* V, const double* coeffs, double sigma, double* D, double* var) {
  BK_TRY(enter(ctx, "bk_deriv_mat"));
  BK_REQUIRE(X && K && V && coeffs && D && var && n > 0 && p > 0 && fits_int(n) && fits_int(p),
             "bk_deriv_mat: bad arguments");
  const int ni = (int)n, pi = (int)p;
  int nbin = 0;
  DevBuf<double> dX, dK, dV, dc, info, W, KW, dD, dR, VR, dvar;
  BK_TRY(h2d(ctx, dX, X, (size_t)n * p));
  BK_TRY(h2d(ctx, dK, K, (size_t)n * n));
  BK_TRY(h2d(ctx, dV, V, (size_t)n * n));
  BK_TRY(h2d(ctx, dc, coeffs, (size_t)n));
  BK_TRY(info.alloc(4 * p + 1));
  BK_TRY(column_binary_info(ctx, dX.p, n, ni, pi, info.p, &nbin));
  const int m = 2 * pi + 2 + 2 * nbin;
  BK_TRY(W.alloc((size_t)n * m));
  BK_TRY(KW.alloc((size_t)n * m));
  BK_TRY(dD.alloc((size_t)n * p));
  BK_TRY(dR.alloc((size_t)n * p));
  BK_TRY(VR.alloc((size_t)n * p));
  BK_TRY(dvar.alloc(p));
  BK_TRY(build_kpass_rhs(ctx, dX.p, n, ni, pi, nbin, dc.p, info.p, W.p, n));
  BK_TRY(gemm(ctx, false, false, ni, m, ni, 1.0, dK.p, n, W.p, n, 0.0, KW.p, n));
  BK_TRY(deriv_epilogue(ctx, dX.p, n, ni, pi, nbin, KW.p, n, info.p, sigma, dD.p, n, dR.p, n));
  BK_TRY(gemm(ctx, false, false, ni, pi, ni, 1.0, dV.p, n, dR.p, n, 0.0, VR.p, n));
  BK_TRY(deriv_variance_dense(ctx, dR.p, n, VR.p, n, ni, pi, info.p, sigma, dvar.p));
  BK_TRY(d2h(ctx, D, dD.p, (size_t)n * p));
  return d2h(ctx, var, dvar.p, (size_t)p);
}

int bk_neffective(bk_ctx* ctx, const double* X, int64_t n, int64_t p, double* out) {
  BK_TRY(enter(ctx, "bk_neffective"));
  BK_REQUIRE(X && out && n > 0 && p > 0 && fits_int(n) && fits_int(p), "bk_neffective: bad arguments");
  DevBuf<double> dX;
  BK_TRY(h2d(ctx, dX, X, (size_t)n * p));
  return neffective_acf(ctx, dX.p, n, (int)n, (int)p, out);
}


int bk_debug_gemm(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                  int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower,
                  int repeats) {
  BK_TRY(enter(ctx, "bk_debug_gemm"));
  BK_REQUIRE(A && B && C && m > 0 && n > 0 && k > 0, "bk_debug_gemm: bad arguments");
  const int64_t a_cols = ta ? m : k, b_cols = tb ? k : n;
  DevBuf<double> dA, dB, dC, dC0;
  BK_TRY(h2d(ctx, dA, A, (size_t)lda * a_cols));
  BK_TRY(h2d(ctx, dB, B, (size_t)ldb * b_cols));
  BK_TRY(h2d(ctx, dC0, C, (size_t)ldc * n));
  BK_TRY(dC.alloc((size_t)ldc * n));
  for (int r = 0; r < std::max(1, repeats); ++r) {
    BK_CUDA(cudaMemcpyAsync(dC.p, dC0.p, sizeof(double) * (size_t)ldc * n, cudaMemcpyDeviceToDevice, ctx->stream));
    BK_TRY(gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, alpha, dA.p, lda, dB.p, ldb, beta, dC.p, ldc,
                lower != 0));
  }
  return d2h(ctx, C, dC.p, (size_t)ldc * n);
}

int bk_debug_sytrd(bk_ctx* ctx, const double* A, int64_t n, double* d, double* e) {
  BK_TRY(enter(ctx, "bk_debug_sytrd"));
  BK_REQUIRE(A && d && e && n > 0 && fits_int(n), "bk_debug_sytrd: bad arguments");
  DevBuf<double> dA, dW, dd, de, dt;
  const long long ldw = sytrd_ld((int)n);
  BK_TRY(h2d(ctx, dA, A, (size_t)n * n));
  BK_TRY(dW.alloc((size_t)ldw * n));
  BK_CUDA(cudaMemsetAsync(dW.p, 0, sizeof(double) * (size_t)ldw * n, ctx->stream));
  BK_TRY(copy_matrix(ctx, dA.p, n, (int)n, (int)n, 1.0, dW.p, ldw));
  BK_TRY(dd.alloc(n));
  BK_TRY(de.alloc(n));
  BK_TRY(dt.alloc(n));
  BK_CUDA(cudaMemsetAsync(de.p, 0, sizeof(double) * n, ctx->stream));
  BK_TRY(sytrd_lower(ctx, dW.p, ldw, (int)n, dd.p, de.p, dt.p, 64, nullptr));
  BK_TRY(d2h(ctx, d, dd.p, n));
  if (n > 1) BK_TRY(d2h(ctx, e, de.p, n - 1));
  return BK_OK;
}

int bk_debug_twostage(bk_ctx* ctx, const double* A, int64_t n, double* band, double* d, double* e, double* Z,
                      int64_t k, double* times) {
  BK_TRY(enter(ctx, "bk_debug_twostage"));
  BK_REQUIRE(A && d && e && n > 0 && fits_int(n) && k >= 0 && k <= n, "bk_debug_twostage: bad arguments");
  DevBuf<double> dA, dd, de, dZ;
  BK_TRY(h2d(ctx, dA, A, (size_t)n * n));
  BK_TRY(dd.alloc(n));
  BK_TRY(de.alloc(n));
  BK_CUDA(cudaMemsetAsync(de.p, 0, sizeof(double) * n, ctx->stream));
  TwoStage ts;
  if (band) {
    // stage 1 only, then copy the band out before stage 2 destroys it
    const int b = sy2sb_bandwidth(), npan = (int)ceil_div(n, b);
    DevBuf<double> w, ab, tst;
    BK_TRY(w.alloc((size_t)n * n));
    BK_TRY(ab.alloc((size_t)2 * b * n));
    BK_TRY(tst.alloc((size_t)npan * b * b));
    BK_TRY(copy_matrix(ctx, dA.p, n, (int)n, (int)n, 1.0, w.p, n));
    BK_TRY(sy2sb(ctx, w.p, n, (int)n, tst.p, ab.p, 2 * b));
    BK_TRY(d2h(ctx, band, ab.p, (size_t)2 * b * n));
  }
  BK_TRY(twostage_reduce(ctx, dA.p, n, (int)n, &ts, dd.p, de.p));
  BK_TRY(d2h(ctx, d, dd.p, n));
  if (n > 1) BK_TRY(d2h(ctx, e, de.p, n - 1));
  if (Z && k > 0) {
    // Z (n x k, host, column-major): in = tridiagonal eigenvectors, out = Q1 Q2 Z
    BK_TRY(h2d(ctx, dZ, Z, (size_t)n * k));
    BK_TRY(twostage_back(ctx, &ts, dZ.p, n, (int)k));
    BK_TRY(d2h(ctx, Z, dZ.p, (size_t)n * k));
  }
  if (times) {
    times[0] = ts.t_sy2sb;
    times[1] = ts.t_sb2st;
    times[2] = ts.t_q2;
    times[3] = ts.t_q1;
  }
  return BK_OK;
}

int bk_debug_stedc(bk_ctx* ctx, const double* d, const double* e, int64_t n, double* evals, double* Z) {
  BK_TRY(enter(ctx, "bk_debug_stedc"));
  BK_REQUIRE(d && e && evals && n > 0 && fits_int(n), "bk_debug_stedc: bad arguments");
  DevBuf<double> dZ;
  if (Z) BK_TRY(dZ.alloc((size_t)n * n));
  int nw = 0;
  BK_TRY(stedc(ctx, (int)n, d, e, evals, Z ? (int)n : 0, -INFINITY, &nw, Z ? dZ.p : nullptr, n, nullptr));
  if (Z) BK_TRY(d2h(ctx, Z, dZ.p, (size_t)n * n));
  return BK_OK;
}

}  // extern "C"
