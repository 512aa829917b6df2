// Symmetric eigensolver internals (device pointers, context stream).
#pragma once
#include "common.cuh"

struct bk_peer;

namespace bk {

// A (n x n, lower triangle referenced and overwritten) -> d[n], e[n-1], tau[n-1] (device);
// reflectors stored LAPACK-style below the first sub-diagonal of A.
struct SytrdStats {
  int launches = 0;
  double kernel_seconds = 0;     // sum of the panel-kernel durations (CUDA events)
  double update_seconds = 0;     // sum of the rank-2nb trailing updates (DMMA GEMM)
  double algorithmic_bytes = 0;  // sum_j 4 (n-1-j)^2: the lower triangle read once per column
};
int sytrd_lower(bk_ctx* ctx, double* A, long long lda, int n, double* d, double* e, double* tau,
                int nb, SytrdStats* stats);

struct StedcStats {
  int levels = 0;
  long long merge_flops = 0;   // flops of the merge GEMMs actually executed
  int top_n = 0, top_k = 0;    // size and non-deflated count of the root merge
};

// Divide & conquer on the tridiagonal (d, e: HOST arrays, n and n-1).  On return
// evals_host[n] ASCENDING.  Eigenvectors: among the max_want largest eigenvalues, those
// with value >= rel_thresh * largest are kept (*n_want of them, the reference's `lastkeeper`
// rule applied to the leading max_want values); Z (n x *n_want, ld = ldz, device, capacity
// n x max_want) gets their eigenvectors, column c <-> the (c+1)-th largest.
int stedc(bk_ctx* ctx, int n, const double* d_host, const double* e_host, double* evals_host,
          int max_want, double rel_thresh, int* n_want, double* Z, long long ldz,
          StedcStats* stats);

// Z (n x k) <- Q Z with Q = H_0 H_1 ... H_{n-2} the reflectors left in A by sytrd_lower.
int ormtr_lower(bk_ctx* ctx, const double* A, long long lda, int n, const double* tau, double* Z,
                long long ldz, int k);

struct BandStats {
  // the two large GEMMs of every stage-1 panel (Z = A22 (V T) and A22 -= [V W][W V]'): launches, summed CUDA-event
  // time of those launches on the library stream, useful flops (the update counts the lower triangle once)
  double gemm_launches = 0, gemm_seconds = 0, gemm_flops = 0;
};

struct EigenTimes {
  double tridiag = 0, dc = 0, backtransform = 0;
  StedcStats dc_stats;
  SytrdStats sytrd;
  int twostage = 0;
  double t_sy2sb = 0, t_sb2st = 0, t_q2 = 0, t_q1 = 0;
  BandStats band;
};

// Full path: K (n x n symmetric, device, preserved) -> evals_host[n] DESCENDING and the
// eigenvectors selected as in stedc().  Works on an internal copy of K whose leading dimension
// is padded to a multiple of 16 doubles (aligned columns for the TMA bulk copies of sytrd).
int eigen_full(bk_ctx* ctx, const double* K, long long ldk, int n, double* evals_host, int max_want,
               double rel_thresh, int* n_want, double* Z, long long ldz, EigenTimes* times);
// Two-stage tridiagonalisation (sy2sb.cu, sb2st.cu): dense -> band (b = 64) -> tridiagonal, and the
// back-transformation Z <- Q1 Q2 Z.  Cheaper than the one-stage reduction when few eigenvectors are wanted:
// stage 1 is GEMM-bound, stage 2 works on the L2-resident band.
int sy2sb_bandwidth();
int sy2sb(bk_ctx* ctx, double* A, long long lda, int n, double* Tstore, double* AB, int ldab,
          BandStats* stats = nullptr, const double* Ksrc = nullptr, long long ldk = 0);
int sb2st(bk_ctx* ctx, double* AB, int n, double* d, double* e, double* VV, double* TAU, int maxhops);
int q2_apply(bk_ctx* ctx, const double* VV, const double* TAU, int maxhops, int n, double* Z, long long ldz, int k);
// GEMM-based variant for many columns (q2_blocked.cu): compact-WY blocks of 64 reflectors, batched DMMA GEMMs
int q2_apply_blocked(bk_ctx* ctx, const double* VV, const double* TAU, int maxhops, int n, double* Z, long long ldz,
                     int k);
int q1_apply(bk_ctx* ctx, const double* A, long long lda, int n, const double* Tstore, double* Z, long long ldz,
             int k);
struct TwoStage {
  DevBuf<double> work, AB, Tstore, VV, TAU;
  int maxhops = 0, n = 0;
  double t_sy2sb = 0, t_sb2st = 0, t_q2 = 0, t_q1 = 0;
  BandStats band;
};
// K (n x n, only read) -> d, e (device, length n); reflectors kept in ts for twostage_back
int twostage_reduce(bk_ctx* ctx, const double* K, long long ldk, int n, TwoStage* ts, double* d, double* e);
int twostage_back(bk_ctx* ctx, TwoStage* ts, double* Z, long long ldz, int k);
// Distributed stage 1 (sy2sb.cu, peer.cuh): the trailing matrix lives block-cyclically on the ranks of `peer`, built
// straight from X (n x p standardised data, Gaussian kernel with bandwidth sigma).  Collective.
size_t sy2sb_dist_heap_bytes(int n);
int sy2sb_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n, double* Afact,
               double* Tstore, double* AB, int ldab, DevBuf<double>& aloc_cache, BandStats* stats);
int twostage_reduce_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n,
                         TwoStage* ts, double* d, double* e);
// Collective full eigensolver on the kernel matrix of X: stage 1 on all ranks, the rest on rank 0.  Only rank 0
// fills evals_host / n_want / Z; the caller broadcasts.
int eigen_full_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n,
                    double* evals_host, int max_want, double rel_thresh, int* n_want, double* Z, long long ldz, size_t off_Q,
                    EigenTimes* times);
static constexpr int kTwoStageFullMax = 16384;  // largest n for which the two-stage path is taken for ALL vectors
bool use_twostage(int n, int max_want, double rel_thresh);
inline long long sytrd_ld(int n) { return ((long long)n + 15) / 16 * 16; }

// Top-k eigenpairs (k << n) by restarted block Krylov + Rayleigh-Ritz (eigen_topk.cu): evals_host[k]
// DESCENDING, Z (n x k) eigenvectors.  K is only read.
struct TopkStats {
  int restarts = 0, matvecs = 0, block = 0, basis = 0;
  double residual = 0;
};
// With `peer`: K is this rank's column block K[:, c0:c0+nloc] and EVERYTHING n-long (basis, K X, residuals) is
// row-partitioned over the ranks; the small projected matrices are all-reduced in a fixed rank order (collective,
// every rank returns the same values and the full set of vectors).
size_t eigen_topk_heap_bytes(int n, int k);
int eigen_topk(bk_ctx* ctx, const double* K, long long ldk, int n, int k, double* evals_host, double* Z,
               long long ldz, TopkStats* stats, bk_peer* peer = nullptr, int c0 = 0, int nloc = 0);
// policy shared by bk_eigen and the fused fit
inline bool use_topk(long long n, long long neig) { return n >= 512 && neig * 3 <= n; }

// Pure host logic of one D&C merge, exported for the CPU unit tests (tests/test_host_logic.py)
struct MergePlan {
  int K = 0;
  double rho = 0;
  std::vector<double> dlam, w;      // non-deflated poles (ascending) and weights
  std::vector<int> nd_cols, nd_type;
  std::vector<int> defl_cols;
  std::vector<double> defl_vals;
  struct Rot { int pj, nj; double c, s; };
  std::vector<Rot> rots;
};
void host_deflate(const double* d, const double* z, int n, int n1, double beta, MergePlan* plan);

}  // namespace bk
