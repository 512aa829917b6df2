// Internal launchers of the non-GEMM kernels (all device pointers, context stream).
#pragma once
#include "common.cuh"

namespace bk {

// ---- gauss_kernel.cu ------------------------------------------------------------------------
int gauss_kernel_sym(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double sigma,
                     double* K, long long ldk);
int gauss_kernel_rect(bk_ctx* ctx, const double* A, long long lda, int m, const double* B,
                      long long ldb, int n, int p, double sigma, double* out, long long ldo);

// ---- elementwise.cu -------------------------------------------------------------------------
// out[:, j] = A[:, j] * (scale ? *scale : 1) * d[j]            (src/multdiag.cpp:17-18)
int col_scale(bk_ctx* ctx, const double* A, long long lda, int n, int k, const double* d,
              const double* dev_scalar, double* out, long long ldo);
// copy the strictly-lower triangle of the n x n matrix onto the upper one
int symmetrize_from_lower(bk_ctx* ctx, double* C, long long ldc, int n);
// d_out[j] = f(ev[j], lambda) for the spectral weights:
//   mode 0: 1/(ev+lam)   mode 1: 1/(ev+lam)^2   mode 2: (ev/(ev+lam))^2
int spectral_weights(bk_ctx* ctx, const double* ev, int k, double lam, int mode, double* out);
// per-column binary detection (src/bigderiv_v3.cpp:28-31): info[4*j+{0,1,2,3}] = {z0, z1, is_binary, slot among
// the binary columns}, info[4p] = number of binary columns (also returned on the host; synchronises the stream).
// info needs 4p+1 doubles.
int column_binary_info(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double* info, int* nbin_host);
// W = [1, c, {x_j or b1_j}, {x_j*c or b1_j*c}, {b0_s}, {b0_s*c}]  (n x (2p+2+2nbin)); for a binary column
// b1_j = [x_j == z1_j], b0_s = [x_j == z0_j] (s = its slot)
int build_kpass_rhs(bk_ctx* ctx, const double* X, long long ldx, int n, int p, int nbin, const double* c,
                    const double* info, double* W, long long ldw);
// From KW = K W:  D (n x p, marginal effects in standardised units) and R (n x p, the vectors
// whose V-quadratic form gives the variance), following src/bigderiv_v3.cpp:31-106 in the
// reduced form of SURVEY.md A.5.
int deriv_epilogue(bk_ctx* ctx, const double* X, long long ldx, int n, int p, int nbin, const double* KW,
                   long long ldkw, const double* info, double sigma, double* D, long long ldd,
                   double* R, long long ldr);
// var[j] = factor_j * sum_i s[i] * G[i,j]^2 ; G = Q'R (k x p); factor from info/sigma/n;
// dev_sigmasq (device scalar) multiplies s when non-null.
int deriv_variance_spectral(bk_ctx* ctx, const double* G, long long ldg, int k, int p,
                            const double* w2, const double* dev_sigmasq, const double* info,
                            double sigma, int n, double* var);
// var[j] = factor_j * sum_i R[i,j] * VR[i,j]   (generic V path of bk_deriv_mat)
int deriv_variance_dense(bk_ctx* ctx, const double* R, long long ldr, const double* VR,
                         long long ldvr, int n, int p, const double* info, double sigma,
                         double* var);
// out[0] = sum_i (y[i]-yhat[i])^2 / n
int residual_sigmasq(bk_ctx* ctx, const double* y, const double* yhat, int n, double* out);
// se2[r] = sum_i s[i] * G[r,i]^2   (diag of G diag(s) G')
int row_quadform(bk_ctx* ctx, const double* G, long long ldg, int m, int k, const double* s,
                 const double* dev_scalar, double host_scale, double* out);
// strided copy: dst (rows x cols, ldd) = alpha * src (rows x cols, lds)
int copy_matrix(bk_ctx* ctx, const double* src, long long lds, int rows, int cols, double alpha,
                double* dst, long long ldd);
// dst[:, j] = src[:, perm[j]]  (column gather), perm on device
int gather_columns(bk_ctx* ctx, const double* src, long long lds, int rows, int cols,
                   const int* perm, double* dst, long long ldd);
int fill(bk_ctx* ctx, double* p, long long n, double v);

// ---- loo.cu ---------------------------------------------------------------------------------
// Leave-one-out loss for nlam candidate lambdas in one pass over Q (n x k, ld = ldq, rows
// [r0, r1) only - a row panel).  z = Q'y (k).  Le_partial[l] = sum_{i in rows} (c_il/d_il)^2.
// If coeffs != nullptr (nlam must be 1) the coefficients c_i of the rows are stored too.
int loo_batch(bk_ctx* ctx, const double* Q, long long ldq, int n_rows, int k, const double* ev,
              const double* z, const double* lambdas_host, int nlam, double* Le_dev,
              double* coeffs);

// ---- neff.cu --------------------------------------------------------------------------------
int neffective_acf(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double* out_host);

}  // namespace bk
