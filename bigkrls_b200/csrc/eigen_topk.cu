// Top-k eigenpairs of a symmetric (PSD kernel) matrix by a restarted block-Krylov method with full
// re-orthogonalisation and Rayleigh-Ritz.  Replaces the `Neig < N` branch of the reference,
// `sp_mat(A)` + `arma::eigs_sym(vals, vecs, sparseA, Neig)` (src/eigen.cpp:18-22; ARPACK/NEWARP implicitly
// restarted Lanczos to machine tolerance) - the dense-to-sparse copy is gone, the dense K is used as is.
// Executable specification: tests/krylov_prototype.py.
//
//   expand   W = K X                          DMMA GEMM  n x n x b      (HBM: one read of K per block)
//            W -= V (V'W)  (twice), X = orth(W)                        (Gram matrix + small eigensolve)
//   extract  H = V'(KV),  (theta, S) = eig(H) (our own dense eigensolver on the m x m matrix)
//            Q = V S, KQ = KV S, residuals ||KQ - Q theta||
//   restart  V <- Q (thick restart), next block = orthonormalised worst residuals
//
// All O(n m b) work is on the GEMM; the host only sees m-vectors (Ritz values, residual norms).
#include <algorithm>
#include <cmath>
#include <numeric>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace bk {

// p[i + j*ld] for local rows i of a matrix whose GLOBAL row index is row0 + i (n_glob rows in total): the value
// depends only on the global position, so a row-partitioned start block is the same matrix for any number of ranks
__global__ void hash_fill_kernel(double* p, int rows, int cols, long long ld, int row0, int n_glob, unsigned long long seed) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % rows), j = (int)(idx / rows);
    unsigned long long x = (unsigned long long)((long long)(row0 + i) + (long long)j * n_glob) * 6364136223846793005ULL + seed;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    p[i + (long long)j * ld] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}

// R[:, j] = KQ[:, j] - theta[j] * Q[:, j];  nrm2[j] = ||R[:, j]||^2 over the local rows   (one CTA per column)
__global__ void ritz_residual_kernel(const double* __restrict__ Q, const double* __restrict__ KQ,
                                     const double* __restrict__ theta, int n, double* __restrict__ R,
                                     double* __restrict__ nrm2) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  const double th = theta[j];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = KQ[i + (long long)j * n] - th * Q[i + (long long)j * n];
    R[i + (long long)j * n] = r;
    s = fma(r, r, s);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) nrm2[j] = s;
}

namespace {

// Everything n-long is ROW-PARTITIONED over the ranks of the peer communicator (nl local rows starting at global
// row c0; single GPU: nl = n): products of the form A'B over the long dimension are local partial sums followed by
// an all-reduce that adds the ranks in a fixed order, so every rank holds the same small matrices bit for bit and
// takes the same decisions; A * (small) products are local.
struct Ws {
  bk_ctx* ctx;
  bk_peer* peer;
  int n, nl, c0;
  size_t stage_off;       // large all-reduce staging in the symmetric heap
  long long slot_elems;
  DevBuf<double> G, Gs, T1, T2, scale;
  int allreduce(double* buf, long long cnt) {
    if (!peer) return BK_OK;
    if (cnt <= 4096) return peer_allreduce_sum(peer, buf, cnt, ctx->stream);
    return peer_allreduce_sum_large(peer, buf, cnt, stage_off, slot_elems, ctx->stream);
  }
};

// X (nl x b local rows, ld nl) <- orthonormal basis of its span via the eigen-decomposition of the Gram matrix,
// applied twice.  Returns the number of columns kept (directions with relative Gram eigenvalue below
// `drop` are discarded - Krylov breakdown / rank deficiency).
int orth_gram(Ws& w, double* X, int b, int* kept) {
  bk_ctx* ctx = w.ctx;
  const int nl = w.nl;
  int r = b;
  for (int pass = 0; pass < 2 && r > 0; ++pass) {
    BK_TRY(w.G.ensure((size_t)r * r));
    BK_TRY(w.Gs.ensure((size_t)r * r));
    BK_TRY(w.T1.ensure((size_t)nl * r));
    BK_TRY(w.scale.ensure(r));
    BK_TRY(gemm(ctx, true, false, r, r, nl, 1.0, X, nl, X, nl, 0.0, w.G.p, r));
    BK_TRY(w.allreduce(w.G.p, (long long)r * r));
    std::vector<double> ev(r);
    int nw = 0;
    BK_TRY(eigen_full(ctx, w.G.p, r, r, ev.data(), r, -INFINITY, &nw, w.Gs.p, r, nullptr));
    int keep = 0;
    std::vector<double> sc(r, 0.0);
    for (int j = 0; j < r; ++j)
      if (ev[j] > 1e-14 * ev[0] && ev[j] > 0.0) {
        sc[j] = 1.0 / std::sqrt(ev[j]);
        keep = j + 1;
      } else {
        break;
      }
    if (keep == 0) {
      r = 0;
      break;
    }
    BK_CUDA(cudaMemcpyAsync(w.scale.p, sc.data(), sizeof(double) * keep, cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(col_scale(ctx, w.Gs.p, r, r, keep, w.scale.p, nullptr, w.Gs.p, r));
    BK_TRY(gemm(ctx, false, false, nl, keep, r, 1.0, X, nl, w.Gs.p, r, 0.0, w.T1.p, nl));
    BK_TRY(copy_matrix(ctx, w.T1.p, nl, nl, keep, 1.0, X, nl));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));  // sc / ev are host temporaries
    r = keep;
  }
  *kept = r;
  return BK_OK;
}

// W (nl x b) -= V (V' W), twice (classical Gram-Schmidt with re-orthogonalisation)
int project_out(Ws& w, const double* V, int m, double* W, int b) {
  if (m <= 0 || b <= 0) return BK_OK;
  BK_TRY(w.T2.ensure((size_t)m * b));
  for (int pass = 0; pass < 2; ++pass) {
    BK_TRY(gemm(w.ctx, true, false, m, b, w.nl, 1.0, V, w.nl, W, w.nl, 0.0, w.T2.p, m));
    BK_TRY(w.allreduce(w.T2.p, (long long)m * b));
    BK_TRY(gemm(w.ctx, false, false, w.nl, b, m, -1.0, V, w.nl, w.T2.p, m, 1.0, W, w.nl));
  }
  return BK_OK;
}

}  // namespace

size_t eigen_topk_heap_bytes(int n, int k) {
  const int b = std::min(128, std::max(8, std::min(n, (k + 3) / 4)));
  const long long m_max = std::min<long long>(n, (long long)k + 8 * b);
  // gathered X block (2 x n x 128), staging of the large all-reduce (2 parities x 8 slots x m_max^2), gathered result
  return sizeof(double) * (2 * (size_t)n * 128 + 2 * 8 * (size_t)(m_max * m_max + 64) + (size_t)n * k) + 16384;
}

int eigen_topk(bk_ctx* ctx, const double* K, long long ldk, int n, int k, double* evals_host, double* Z,
               long long ldz, TopkStats* stats, bk_peer* peer, int c0, int nloc) {
  BK_REQUIRE(k >= 1 && k <= n, "eigen_topk: k must be in 1..n");
  const int b = std::min(128, std::max(8, std::min(n, (k + 3) / 4)));
  const int m_max = std::min(n, k + 8 * b);
  const double tol = 2e-13;
  const int nl = peer ? nloc : n;      // local rows
  const int r0 = peer ? c0 : 0;
  Ws w;
  w.ctx = ctx;
  w.peer = peer;
  w.n = n;
  w.nl = nl;
  w.c0 = r0;
  w.stage_off = 0;
  w.slot_elems = (long long)m_max * m_max + 64;
  // K X.  Single GPU: one GEMM over the whole K.  With a peer communicator K stays PARTITIONED (`K` is this rank's
  // column block K[:, c0:c0+nloc] = its row panel, K is symmetric): the rows of X are gathered from all ranks into a
  // symmetric buffer by peer stores over NVLink, and K[:, own]' X gives this rank's rows of K X.
  size_t x_off[2] = {0, 0}, z_off = 0;
  unsigned mv_count = 0;
  if (peer) {
    for (int q = 0; q < 2; ++q) BK_TRY(peer_alloc(peer, sizeof(double) * (size_t)n * 128, &x_off[q]));
    BK_TRY(peer_alloc(peer, sizeof(double) * 2 * (size_t)peer->world * (size_t)w.slot_elems, &w.stage_off));
    BK_TRY(peer_alloc(peer, sizeof(double) * (size_t)n * k, &z_off));
    BK_TRY(peer_barrier(peer, ctx->stream));
  }
  auto matvec = [&](const double* X, int bx, double* KX) -> int {
    if (!peer) return gemm(ctx, false, false, n, bx, n, 1.0, K, ldk, X, n, 0.0, KX, n);
    const int q = (int)(mv_count++ & 1u);
    const unsigned all = (1u << peer->world) - 1u;
    const unsigned seq = peer_next_seq(peer, CH_KRYLOV);
    BK_TRY(peer_push2d(peer, X, nl, nl, bx, x_off[q] + sizeof(double) * (size_t)r0, n, all, CH_KRYLOV, seq, ctx->stream));
    BK_TRY(peer_wait(peer, CH_KRYLOV, all, seq, ctx->stream));
    return gemm(ctx, true, false, nl, bx, n, 1.0, K, ldk, peer_ptr(peer, x_off[q]), n, 0.0, KX, nl);
  };
  DevBuf<double> V, KV, Wt, Qb, KQb, H, S, theta_d, nrm_d, Rb;
  DevBuf<int> idx_d;
  BK_TRY(V.alloc((size_t)nl * (m_max + b)));
  BK_TRY(KV.alloc((size_t)nl * (m_max + b)));
  BK_TRY(Wt.alloc((size_t)nl * b));
  BK_TRY(Qb.alloc((size_t)nl * k));
  BK_TRY(KQb.alloc((size_t)nl * k));
  BK_TRY(Rb.alloc((size_t)nl * k));
  BK_TRY(H.alloc((size_t)m_max * m_max));
  BK_TRY(S.alloc((size_t)m_max * k));
  BK_TRY(theta_d.alloc(k));
  BK_TRY(nrm_d.alloc(k));
  BK_TRY(idx_d.alloc(b));

  // deterministic pseudo-random start block (a function of the global position only)
  hash_fill_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(V.p, nl, b, nl, r0, n, 0x9E3779B97F4A7C15ULL);
  BK_LAUNCHED(ctx);
  int bx = 0;
  BK_TRY(orth_gram(w, V.p, b, &bx));
  int m = 0;  // accepted basis columns; the candidate block X sits at V[:, m : m+bx]
  std::vector<double> th(k), nrm(k);
  int outer = 0, matvecs = 0;
  const int max_outer = 300;
  for (; outer < max_outer; ++outer) {
    // ---- expand the Krylov basis ------------------------------------------------------------------
    while (bx > 0 && m + bx <= m_max) {
      double* X = V.p + (size_t)m * nl;
      double* KX = KV.p + (size_t)m * nl;
      BK_TRY(matvec(X, bx, KX));
      matvecs += bx;
      m += bx;
      BK_TRY(copy_matrix(ctx, KX, nl, nl, bx, 1.0, Wt.p, nl));
      BK_TRY(project_out(w, V.p, m, Wt.p, bx));
      int r = 0;
      BK_TRY(orth_gram(w, Wt.p, bx, &r));
      bx = r;
      if (bx > 0 && m + bx <= m_max + b) BK_TRY(copy_matrix(ctx, Wt.p, nl, nl, bx, 1.0, V.p + (size_t)m * nl, nl));
      if (m + bx > m_max) break;
    }
    if (m < k) {
      // the Krylov space is exhausted (exact invariant subspace smaller than k): restart direction
      set_error("eigen_topk: Krylov breakdown with %d < %d basis vectors", m, k);
      return BK_ERR_NUMERIC;
    }
    // ---- Rayleigh-Ritz -----------------------------------------------------------------------------
    BK_TRY(gemm(ctx, true, false, m, m, nl, 1.0, V.p, nl, KV.p, nl, 0.0, H.p, m));
    BK_TRY(w.allreduce(H.p, (long long)m * m));
    std::vector<double> ev(m);
    int nw = 0;
    BK_TRY(eigen_full(ctx, H.p, m, m, ev.data(), k, -INFINITY, &nw, S.p, m, nullptr));
    for (int j = 0; j < k; ++j) th[j] = ev[j];
    BK_CUDA(cudaMemcpyAsync(theta_d.p, th.data(), sizeof(double) * k, cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(gemm(ctx, false, false, nl, k, m, 1.0, V.p, nl, S.p, m, 0.0, Qb.p, nl));
    BK_TRY(gemm(ctx, false, false, nl, k, m, 1.0, KV.p, nl, S.p, m, 0.0, KQb.p, nl));
    ritz_residual_kernel<<<k, 256, 0, ctx->stream>>>(Qb.p, KQb.p, theta_d.p, nl, Rb.p, nrm_d.p);
    BK_LAUNCHED(ctx);
    BK_TRY(w.allreduce(nrm_d.p, k));
    BK_CUDA(cudaMemcpyAsync(nrm.data(), nrm_d.p, sizeof(double) * k, cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    double worst = 0.0;
    for (int j = 0; j < k; ++j) worst = std::max(worst, std::sqrt(nrm[j]));
    if (!(worst == worst)) {
      set_error("eigen_topk: non-finite residual (NaN/Inf in the input?)");
      return BK_ERR_NUMERIC;
    }
    if (worst <= tol * std::fabs(th[0]) || m >= n) break;
    // ---- thick restart with the Ritz vectors; next block = worst residual directions ---------------
    std::vector<int> order(k);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return nrm[a] > nrm[c]; });
    const int nb = std::min(b, k);
    BK_CUDA(cudaMemcpyAsync(idx_d.p, order.data(), sizeof(int) * nb, cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(gather_columns(ctx, Rb.p, nl, nl, nb, idx_d.p, Wt.p, nl));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    BK_TRY(copy_matrix(ctx, Qb.p, nl, nl, k, 1.0, V.p, nl));
    BK_TRY(copy_matrix(ctx, KQb.p, nl, nl, k, 1.0, KV.p, nl));
    m = k;
    BK_TRY(project_out(w, V.p, m, Wt.p, nb));
    BK_TRY(orth_gram(w, Wt.p, nb, &bx));
    if (bx > 0) BK_TRY(copy_matrix(ctx, Wt.p, nl, nl, bx, 1.0, V.p + (size_t)m * nl, nl));
    if (bx == 0) break;  // residuals vanish numerically: converged as far as FP64 goes
  }
  double worst = 0.0;
  for (int j = 0; j < k; ++j) worst = std::max(worst, std::sqrt(nrm[j]));
  if (worst > 1e-9 * std::fabs(th[0])) {
    set_error("eigen_topk: no convergence after %d restarts (residual %.3e relative)", outer,
              worst / std::fabs(th[0]));
    return BK_ERR_NUMERIC;
  }
  for (int j = 0; j < k; ++j) evals_host[j] = th[j];
  if (Z) {
    if (!peer) {
      BK_TRY(copy_matrix(ctx, Qb.p, n, n, k, 1.0, Z, ldz));
    } else {
      // every rank needs all rows of the eigenvectors: all-gather of the row panels by peer stores
      const unsigned all = (1u << peer->world) - 1u;
      const unsigned seq = peer_next_seq(peer, CH_KRYLOV);
      BK_TRY(peer_push2d(peer, Qb.p, nl, nl, k, z_off + sizeof(double) * (size_t)r0, n, all, CH_KRYLOV, seq, ctx->stream));
      BK_TRY(peer_wait(peer, CH_KRYLOV, all, seq, ctx->stream));
      BK_TRY(copy_matrix(ctx, peer_ptr(peer, z_off), n, n, k, 1.0, Z, ldz));
    }
  }
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (peer) {
    BK_TRY(peer_barrier(peer, ctx->stream));
    BK_TRY(peer_check(peer, ctx->stream));
  }
  if (stats) {
    stats->restarts = outer + 1;
    stats->matvecs = matvecs;
    stats->block = b;
    stats->basis = m_max;
    stats->residual = worst / std::fabs(th[0]);
  }
  return BK_OK;
}

}  // namespace bk
