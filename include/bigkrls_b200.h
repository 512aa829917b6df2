/*
 * bigkrls_b200 — C ABI of the B200-native bigKRLS estimation hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no R, no
 * torch types.  All matrices are FP64, COLUMN-MAJOR (R / bigmemory layout) with no padding
 * unless a leading dimension is given.  Every function returns 0 on success or a negative
 * bk_status; bk_last_error() returns a human-readable message for the calling thread.
 * There is NO CPU fallback: every entry point that computes needs a CUDA device and fails
 * with BK_ERR_CUDA otherwise.
 *
 * Two families of entry points:
 *
 *  (1) per-op functions on HOST buffers, one per native export of the reference
 *      (reference: R/RcppExports.R:4-46, registered src/RcppExports.cpp:147-160).  They
 *      copy in, compute on the GPU, copy out - used by the `.Call` shim for users who call
 *      bigKRLS:::bGaussKernel / bEigen / bSolveForc directly, and by the stage-level parity
 *      tests.
 *
 *  (2) a fused, handle-based fit (bk_fit_*) that keeps X, K, Q, Lambda on the device between
 *      stages; this is what bigKRLS(), predict() and crossvalidate.bigKRLS() call
 *      (reference driver: R/bigKRLS.R:262-329).  Row(-panel)-partitioned multi-GPU operation
 *      is expressed through a small communicator vtable (bk_comm) whose callbacks the host
 *      (torch.distributed/NCCL in the Python host, or anything else) provides.
 */
#ifndef BIGKRLS_B200_H
#define BIGKRLS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BK_VERSION 100

#if defined(__GNUC__)
#define BK_API __attribute__((visibility("default")))
#else
#define BK_API
#endif

typedef enum {
  BK_OK = 0,
  BK_ERR_ARG = -1,      /* bad argument (null pointer, non-positive size, ...) */
  BK_ERR_CUDA = -2,     /* CUDA runtime error / no device */
  BK_ERR_NUMERIC = -3,  /* numerical failure (NaN eigenvalues, no convergence) */
  BK_ERR_STATE = -4,    /* stage called out of order */
  BK_ERR_COMM = -5      /* communicator callback failed */
} bk_status;

typedef struct bk_ctx bk_ctx;   /* one per process x device */
typedef struct bk_fit bk_fit;   /* one fitted (or being fitted) model, device resident */

/* ---- context ------------------------------------------------------------------------ */
BK_API int bk_version(void);
BK_API const char* bk_last_error(void);
BK_API int bk_init(int device, bk_ctx** out);
BK_API void bk_destroy(bk_ctx* ctx);
/* Gives cached device memory back: the eigensolver's work matrices kept in the context between fits and the
 * unused part of the device memory pool.  Results of live fits are not affected. */
BK_API int bk_trim(bk_ctx* ctx);
/* name, SM count, total/free HBM bytes of the context's device */
BK_API int bk_device_info(bk_ctx* ctx, char* name, int name_len, int* sm_count, int64_t* hbm_total,
                   int64_t* hbm_free);
/* pinned host memory for the caller's staging buffers (D2H of the N x N outputs) */
/* number of kernels this context has launched so far (bench.py's "gpu_launches") */
BK_API int64_t bk_launch_count(bk_ctx* ctx);
BK_API int bk_host_alloc(bk_ctx* ctx, int64_t bytes, void** out);
BK_API int bk_host_free(bk_ctx* ctx, void* p);

/* ---- (1) per-op entry points, HOST pointers ------------------------------------------- */

/* BigGaussKernel(pA, pOut, sigma)            src/gauss_kernel.cpp:14-42
 * K[i,j] = exp(-sum_d (X[i,d]-X[j,d])^2 / sigma);  X n x p, K n x n (caller allocated). */
BK_API int bk_gauss_kernel(bk_ctx* ctx, const double* X, int64_t n, int64_t p, double sigma, double* K);

/* BigTempKernel(pA, pB, pOut, sigma)         src/temp_kernel.cpp:14-44
 * out[i,j] = exp(-||A_i - B_j||^2 / sigma);  A m x p, B n x p, out m x n. */
BK_API int bk_temp_kernel(bk_ctx* ctx, const double* A, int64_t m, const double* B, int64_t n, int64_t p,
                   double sigma, double* out);

/* BigEigen(pA, Neig, pValBigMat, pVecBigMat) src/eigen.cpp:14-45
 * Symmetric A n x n -> vals (neig, DESCENDING) and vecs n x neig.  neig == n: full
 * decomposition (eig_sym); neig < n: the neig largest (eigs_sym).  vecs may be NULL
 * (values only).  Signs of eigenvectors are arbitrary, as in LAPACK. */
BK_API int bk_eigen(bk_ctx* ctx, const double* A, int64_t n, int64_t neig, double* vals, double* vecs);

/* BigSolveForc(pEigenvectors, Eigenvalues, y, lambda) -> list(Le, coeffs)
 *                                            src/solveforc.cpp:14-78
 * Q n x k, evals: at least k values (only the first k are used). */
BK_API int bk_solve_for_c(bk_ctx* ctx, const double* Q, int64_t n, int64_t k, const double* evals,
                   const double* y, double lambda, double* Le, double* coeffs);

/* Batched LOO loss: Le[l] for nlam candidate lambdas in ONE pass over Q (north_star 3). */
BK_API int bk_loo_batch(bk_ctx* ctx, const double* Q, int64_t n, int64_t k, const double* evals,
                 const double* y, const double* lambdas, int nlam, double* Le);

/* BigMultDiag(pA, diag, pOut)                src/multdiag.cpp:14-37   out[:,i] = A[:,i]*diag[i] */
BK_API int bk_mult_diag(bk_ctx* ctx, const double* A, int64_t n, int64_t k, const double* diag,
                 double* out);

/* BigCrossProd(pA, pB, pOut)  out = A'B      src/crossprod.cpp:13-30  A r x m, B r x n, out m x n */
BK_API int bk_crossprod(bk_ctx* ctx, const double* A, int64_t r, int64_t m, const double* B, int64_t n,
                 double* out);
/* BigXtX(pA, pOut)            out = A'A      src/crossprod.cpp:32-48 */
BK_API int bk_xtx(bk_ctx* ctx, const double* A, int64_t r, int64_t m, double* out);
/* BigTCrossProd(pA, pB, pOut) out = A B'     src/crossprod.cpp:50-68  A m x r, B n x r, out m x n */
BK_API int bk_tcrossprod(bk_ctx* ctx, const double* A, int64_t m, int64_t r, const double* B, int64_t n,
                  double* out);
/* BigXXt(pA, pOut)            out = A A'     src/crossprod.cpp:70-85 */
BK_API int bk_xxt(bk_ctx* ctx, const double* A, int64_t m, int64_t r, double* out);

/* BigDerivMat(pX, pK, pVCovMatC, pDerivatives, pVarAvgDerivatives, coeffs, sigma)
 *                                            src/bigderiv_v3.cpp:14-132
 * X n x p (standardised columns to differentiate), K n x n, V n x n, coeffs n.
 * D n x p, var p.  Binary columns (exactly two distinct values) get first differences. */
BK_API int bk_deriv_mat(bk_ctx* ctx, const double* X, int64_t n, int64_t p, const double* K,
                 const double* V, const double* coeffs, double sigma, double* D, double* var);

/* BigNeffective(pX) -> double                src/Neffective.cpp:14-76 */
BK_API int bk_neffective(bk_ctx* ctx, const double* X, int64_t n, int64_t p, double* out);

/* General FP64 GEMM on host buffers (bigalgebra `%*%` role, R/bigKRLS.R:291,307,601):
 * C (m x n) = op(A) op(B); ta/tb != 0 means transposed operand. */
BK_API int bk_dgemm(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, const double* A,
             int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc);

/* ---- (2) fused fit --------------------------------------------------------------------- */

/* Communicator for row-partitioned multi-GPU runs: one process per GPU.  All buffers are
 * DEVICE pointers on the calling rank's GPU; the callbacks must be complete (stream-
 * synchronised) when they return.  world == 1 -> every callback may be NULL. */
typedef struct bk_peer bk_peer;   /* native peer-memory communicator (NVLink, CUDA IPC), see bk_peer_create */
typedef struct bk_comm {
  int rank;
  int world;
  void* user;
  /* in-place sum over ranks of n doubles */
  int (*allreduce_sum)(void* user, double* dev_buf, int64_t n);
  /* rank r contributes counts[r] doubles found at dev_buf + displs[r]; afterwards every rank
   * holds all segments (in-place all-gather-v) */
  int (*allgatherv)(void* user, double* dev_buf, const int64_t* counts, const int64_t* displs);
  /* broadcast n doubles from `root` */
  int (*broadcast)(void* user, double* dev_buf, int64_t n, int root);
  /* optional (may be NULL): when set, the fit does not use the callbacks above at all - every exchange is a
   * kernel of the library storing into the other ranks' HBM over NVLink (no host synchronisation), the dense->band
   * stage of the eigensolver is distributed over the ranks, and the block-Krylov path keeps K partitioned. */
  bk_peer* peer;
} bk_comm;

/* Bootstrap for bk_peer_create: a HOST all-gather of `bytes` bytes per rank (recv holds world * bytes, in rank
 * order) - torch.distributed.all_gather_object, MPI_Allgather, a socket ...  Called a handful of times when the
 * communicator is created or its symmetric heap grows, never on the data path. */
typedef int (*bk_exchange_fn)(void* user, const void* send, void* recv, int64_t bytes);
/* Collective over all ranks (one process per GPU, at most 8, one node with NVLink peer access): every rank
 * allocates a symmetric heap, the CUDA IPC handles travel through `exchange`, and every heap is mapped into every
 * process. */
BK_API int bk_peer_create(bk_ctx* ctx, int rank, int world, bk_exchange_fn exchange, void* user, bk_peer** out);
BK_API void bk_peer_destroy(bk_peer* p);   /* collective */
/* collective self-test of the peer collectives (all-reduce, all-gather-v, broadcast): *failures = mismatches */
BK_API int bk_peer_selftest(bk_peer* p, int* failures);

typedef struct bk_fit_opts {
  double sigma;            /* kernel bandwidth (R default: ncol(X)) */
  double eigtrunc;         /* keep eigenpairs with value >= eigtrunc * largest (R/bigKRLS_Rcpp_functions.R:190) */
  int64_t neig;            /* number of eigenpairs (n = full decomposition) */
  double lambda;           /* > 0: user lambda, skip the search; <= 0: golden-section search */
  double L, U;             /* search bounds; L < 0 (or NaN) / U <= 0: not supplied -> the reference's bounds loops
                              (a user L = 0 is legal, R/bigKRLS.R:225-228) */
  double tol;              /* <= 0: 1e-3 * n (what the reference always ends up using) */
  int derivative;          /* compute marginal effects */
  int vcov;                /* compute vcov.c / vcov.fitted */
  int n_which;             /* number of derivative columns; 0 = all */
  const int32_t* which;    /* 0-based column indices, n_which of them */
  double y_sd;             /* sd(y) of the un-standardised y: folded into vcov outputs (R/bigKRLS.R:439,446) */
  int loo_batch;           /* lambda candidates evaluated per pass over Q (speculative tree), 1..15; 0 = default 15 */
  int keep_vcov_fitted;    /* 0: skip vcov.fitted (e.g. folds of a cross-validation) */
  double* K_host;          /* optional HOST destination for this rank's column block of K (n x (c1-c0), c0 = n*rank/world,
                              c1 = n*(rank+1)/world): the copy is queued on a second stream as soon as the kernel stage
                              is done and overlaps the eigensolver; complete when bk_fit_run returns (pinned memory
                              recommended).  bk_fit_get_K with the same pointer is then a no-op. */
} bk_fit_opts;

BK_API void bk_fit_default_opts(bk_fit_opts* o, int64_t n, int64_t p);

/* Summary scalars of a fit. */
typedef struct bk_fit_info {
  int64_t n, p, neig, lastkeeper, n_deriv;
  double lambda, Le, sigmasq, neffective;
  int n_probes;            /* LOO evaluations consumed by the golden-section search */
  int n_passes;            /* passes over Q made for them (batched) */
  /* device seconds per stage (CUDA events on the library stream) */
  double t_kernel, t_eigen, t_lambda, t_coef, t_vcov, t_deriv, t_total;
  /* eigensolver break-down */
  double t_tridiag, t_dc, t_backtransform;
  /* dominant kernel (tridiagonalisation panel kernel): launches, summed device time of those
   * launches (CUDA events on the launching stream) and their algorithmic HBM bytes */
  double sytrd_launches, sytrd_kernel_seconds, sytrd_bytes;
  /* divide & conquer: tree levels, flops of the merge GEMMs, size / non-deflated count of the root merge */
  double dc_levels, dc_merge_flops, dc_top_n, dc_top_k;
  /* kernels launched by the library during this fit */
  double gpu_launches;
  /* Neig << N path (block Krylov): matrix-vector products with K and thick restarts (0 on the full path) */
  double krylov_matvecs, krylov_restarts;
  /* two-stage tridiagonalisation (taken when few eigenvectors are wanted): 1 if used, device seconds of
   * dense->band, band->tridiagonal, and the two back-transformations */
  double twostage, t_sy2sb, t_sb2st, t_q2, t_q1;
  /* its dominant kernel (the FP64 DMMA GEMM of the stage-1 panel updates): launches, summed CUDA-event time of
   * those launches on the library stream, useful flops */
  double band_gemm_launches, band_gemm_seconds, band_gemm_flops;
} bk_fit_info;

/* Xs (n x p) and ys (n) are the STANDARDISED data (R/bigKRLS.R:251-254), host pointers.
 * Runs all five stages on the device.  comm may be NULL (single GPU). */
BK_API int bk_fit_run(bk_ctx* ctx, const double* Xs, const double* ys, int64_t n, int64_t p,
               const bk_fit_opts* opts, const bk_comm* comm, bk_fit** out);
/* Same, inputs already resident on the device (device pointers). */
BK_API int bk_fit_run_device(bk_ctx* ctx, const double* dXs, const double* dys, int64_t n, int64_t p,
                      const bk_fit_opts* opts, const bk_comm* comm, bk_fit** out);
BK_API void bk_fit_free(bk_fit* f);
BK_API int bk_fit_get_info(const bk_fit* f, bk_fit_info* info);

/* Getters: copy a result field to a HOST buffer.  N x N fields under a communicator are
 * column blocks: rank r owns columns [bk_fit_col0(r), bk_fit_col1(r)) and the getter writes
 * only that block (n x ncols, contiguous) - the caller places it (symmetric matrices:
 * column block == transposed row panel). */
BK_API int bk_fit_col_range(const bk_fit* f, int64_t* c0, int64_t* c1);
BK_API int bk_fit_get_K(const bk_fit* f, double* host);               /* kernel, n x (c1-c0) */
BK_API int bk_fit_get_eigenvalues(const bk_fit* f, double* host);     /* neig, descending */
BK_API int bk_fit_get_eigenvectors(const bk_fit* f, double* host);    /* n x lastkeeper */
BK_API int bk_fit_get_coeffs(const bk_fit* f, double* host);          /* n */
BK_API int bk_fit_get_yfitted(const bk_fit* f, double* host);         /* n, standardised units */
BK_API int bk_fit_get_vcov_c(const bk_fit* f, double* host);          /* y_sd^2 * vcov(c), n x (c1-c0) */
BK_API int bk_fit_get_vcov_fitted(const bk_fit* f, double* host);     /* y_sd^2 * vcov(yhat), n x (c1-c0) */
BK_API int bk_fit_get_derivatives(const bk_fit* f, double* host);     /* n x n_deriv, standardised units (all rows) */
BK_API int bk_fit_get_var_avgderiv(const bk_fit* f, double* host);    /* n_deriv */
BK_API int bk_fit_get_binary(const bk_fit* f, int32_t* host);         /* n_deriv flags: column got first differences */

/* predict.bigKRLS (R/bigKRLS.R:590-616): newXs m x p standardised with the TRAINING
 * mean/sd.  pred_std (m) = Knew c.  Knew (m x n) and se2 (m, diag of Knew V Knew', in
 * y_sd^2 units, before the Neffective correction) may be NULL. */
BK_API int bk_fit_predict(const bk_fit* f, const double* newXs, int64_t m, double* pred_std,
                   double* Knew, double* se2);
/* Same, and vcov_pred (m x m, may be NULL) = Knew vcov.est.c Knew' - the reference's `vcov.est.pred`
 * (R/bigKRLS.R:605; y_sd^2 units, before the Neffective correction), formed spectrally as
 * y_sd^2 sigmasq (Knew Q) diag((ev+lambda)^-2) (Knew Q)'. */
BK_API int bk_fit_predict_full(const bk_fit* f, const double* newXs, int64_t m, double* pred_std,
                        double* Knew, double* se2, double* vcov_pred);

/* ---- test / measurement hooks ---------------------------------------------------------- */
/* FP64 micro-benchmarks used to establish the roofline denominators (tools/):
 * kind 0: DFMA issue-bound loop, 1: DMMA m8n8k4 loop, 2: HBM copy.  Returns TFLOP/s or GB/s.
 * kind 7: latency in cycles of a dependent chain (size = 0 DFMA, 1 DMMA, 2 DADD, 3 shuffle + DADD, 4 shared-memory round trip). */
BK_API int bk_microbench(bk_ctx* ctx, int kind, int64_t size, int iters, double* result);
/* Time (device seconds, CUDA events) of `iters` runs of the library DGEMM on device-
 * generated data: C(m x n) = alpha op(A) op(B) + beta C; lower = 1 computes only the tiles on/below the
 * diagonal, lower = 2 additionally mirrors them into the upper triangle. */
BK_API int bk_dgemm_bench(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, int lower, double beta,
                   int iters, double* seconds);

/* Stage-level hooks of the eigensolver (tests/test_gpu_ops.py, tests/test_gpu_twostage.py): tridiagonalisation only
 * (A n x n host -> d[n], e[n-1]) and divide & conquer only (d, e host -> evals ascending, Z n x n
 * eigenvectors of the tridiagonal, column c <-> c-th LARGEST; Z may be NULL). */
/* general library GEMM on host buffers, C (in/out) = alpha op(A) op(B) + beta C, optional lower-tile mode */
BK_API int bk_debug_gemm(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, double alpha, const double* A,
                  int64_t lda, const double* B, int64_t ldb, double beta, double* C, int64_t ldc, int lower,
                  int repeats);
BK_API int bk_debug_sytrd(bk_ctx* ctx, const double* A, int64_t n, double* d, double* e);
/* Two-stage reduction (dense -> band -> tridiagonal).  band (optional, 128 x n) receives the stage-1 band in
 * lower storage band[d + j*128] = B[j+d, j]; d/e the tridiagonal; Z (optional, n x k): on entry eigenvectors of
 * the tridiagonal, on exit Q1 Q2 Z; times (optional, 4): seconds of sy2sb, sb2st, Q2, Q1. */
BK_API int bk_debug_twostage(bk_ctx* ctx, const double* A, int64_t n, double* band, double* d, double* e, double* Z,
                             int64_t k, double* times);
BK_API int bk_debug_stedc(bk_ctx* ctx, const double* d, const double* e, int64_t n, double* evals, double* Z);

/* ---- host-logic hooks (pure CPU, no device needed) ------------------------------------------
 * The golden-section search with speculative batching (fit.cu) and the divide-and-conquer
 * deflation bookkeeping (stedc.cu) are host code; these two entry points expose them so the
 * CPU test-suite can check them against the oracle / the numpy prototype without a GPU. */
typedef int (*bk_le_callback)(void* user, const double* lambdas, int nlam, double* Le_out);
BK_API int bk_host_lambda_search(const double* evals, int64_t neig, int64_t n, double L, double U, double tol,
                          int batch, bk_le_callback cb, void* user, double* lambda, double* L_out,
                          double* U_out, int* probes, int* passes);
BK_API int bk_host_deflate_test(const double* d, const double* z, int n, int n1, double beta, int* K,
                         double* dlam, double* w, int32_t* nd_cols, int32_t* nd_type,
                         int32_t* defl_cols, double* defl_vals, int* nrot, int32_t* rot_idx,
                         double* rot_cs);

#ifdef __cplusplus
}
#endif
#endif /* BIGKRLS_B200_H */
