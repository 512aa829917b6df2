// Back-transformation of the tridiagonal eigenvectors, Z <- Q Z with Q = H_0 H_1 ... H_{n-2}
// the Householder reflectors left below the sub-diagonal by sytrd.cu (LAPACK dormtr role inside
// dsyevd; reference src/eigen.cpp:24), plus the full-eigensolver driver.
//
// Blocked compact-WY form: for each block of ib reflectors (last block first)
//     Z[j0+1:, :] -= V (T (V' Z[j0+1:, :]))        H_j0 ... H_{j0+ib-1} = I - V T V'
// V is unpacked to an explicit unit-lower-trapezoidal panel so that all three products run
// on the DMMA GEMM; T comes from S = V'V (split-K GEMM) and an ib-step triangular recurrence.
// Only the k = lastkeeper columns that the fit will use are transformed: 4 n k ib flops per
// block, 2 n^2 k in total.
#include <cstdlib>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace bk {

__global__ void unpack_v_kernel(const double* __restrict__ A, long long lda, int j0, int ib,
                                int mrows, double* __restrict__ Vb) {
  const long long total = (long long)mrows * ib;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % mrows), c = (int)(idx / mrows);
    double v;
    if (r < c)
      v = 0.0;
    else if (r == c)
      v = 1.0;
    else
      v = A[(j0 + 1 + r) + (long long)(j0 + c) * lda];
    Vb[r + (long long)c * mrows] = v;
  }
}

// T (ib x ib upper triangular) from S = V'V and tau:  T[i,i] = tau_i,
// T[0:i, i] = -tau_i * T[0:i,0:i] * S[0:i, i]
__global__ void larft_kernel(const double* __restrict__ S, const double* __restrict__ tau, int ib,
                             double* __restrict__ T) {
  extern __shared__ double ts[];  // ib x ib
  const int r = threadIdx.x;
  for (int idx = threadIdx.x; idx < ib * ib; idx += blockDim.x) ts[idx] = 0.0;
  __syncthreads();
  for (int i = 0; i < ib; ++i) {
    const double ti = tau[i];
    double acc = 0.0;
    if (r < i) {
      for (int q = r; q < i; ++q) acc = fma(ts[r + q * ib], S[q + (long long)i * ib], acc);
    }
    __syncthreads();
    if (r < i) ts[r + i * ib] = -ti * acc;
    if (r == i) ts[i + i * ib] = ti;
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < ib * ib; idx += blockDim.x) T[idx] = ts[idx];
}

int ormtr_lower(bk_ctx* ctx, const double* A, long long lda, int n, const double* tau, double* Z,
                long long ldz, int k) {
  if (n <= 1 || k <= 0) return BK_OK;
  const int IB = 128;  // wider blocks halve the passes over Z (the k = ib GEMMs are traffic bound)
  const int nref = n - 1;  // reflectors live in columns 0 .. n-2
  DevBuf<double> Vb, S, T, W1, W2;
  BK_TRY(Vb.alloc((size_t)n * IB));
  BK_TRY(S.alloc((size_t)IB * IB));
  BK_TRY(T.alloc((size_t)IB * IB));
  BK_TRY(W1.alloc((size_t)IB * k));
  BK_TRY(W2.alloc((size_t)IB * k));
  BK_CUDA(cudaFuncSetAttribute(larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(double) * IB * IB)));
  const int nblocks = (int)ceil_div(nref, IB);
  for (int b = nblocks - 1; b >= 0; --b) {
    const int j0 = b * IB;
    const int ib = std::min(IB, nref - j0);
    const int mrows = n - j0 - 1;
    const long long tot = (long long)mrows * ib;
    unpack_v_kernel<<<(unsigned)std::min<long long>(ceil_div(tot, 256), 8LL * ctx->sm_count), 256, 0,
                      ctx->stream>>>(A, lda, j0, ib, mrows, Vb.p);
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, true, false, ib, ib, mrows, 1.0, Vb.p, mrows, Vb.p, mrows, 0.0, S.p, ib));
    larft_kernel<<<1, IB, sizeof(double) * ib * ib, ctx->stream>>>(S.p, tau + j0, ib, T.p);
    BK_LAUNCHED(ctx);
    BK_CUDA(cudaGetLastError());
    double* Zb = Z + (j0 + 1);
    BK_TRY(gemm(ctx, true, false, ib, k, mrows, 1.0, Vb.p, mrows, Zb, ldz, 0.0, W1.p, ib));
    BK_TRY(gemm(ctx, false, false, ib, k, ib, 1.0, T.p, ib, W1.p, ib, 0.0, W2.p, ib));
    BK_TRY(gemm(ctx, false, false, mrows, k, ib, -1.0, Vb.p, mrows, W2.p, ib, 1.0, Zb, ldz));
  }
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

// *too_wide is set (and nothing is written) when the eigenvalue threshold keeps so many eigenvectors that the
// one-stage reduction with its GEMM-bound back-transformation is the cheaper way to get them.
static int eigen_full_twostage(bk_ctx* ctx, const double* K, long long ldk, int n, double* evals_host,
                               int max_want, double rel_thresh, int* n_want, double* Z, long long ldz,
                               EigenTimes* times, bool* too_wide) {
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  DevBuf<double> d, e;
  BK_TRY(d.alloc(n));
  BK_TRY(e.alloc(n));
  BK_CUDA(cudaMemsetAsync(e.p, 0, sizeof(double) * n, ctx->stream));
  TwoStage ts;
  tm.start();
  BK_TRY(twostage_reduce(ctx, K, ldk, n, &ts, d.p, e.p));
  const double t_tri = tm.stop();
  std::vector<double> dh(n), eh(n), ev(n);
  BK_CUDA(cudaMemcpyAsync(dh.data(), d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaMemcpyAsync(eh.data(), e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(dh[i]) || !std::isfinite(eh[i])) {
      set_error("eigen: tridiagonalisation produced a non-finite entry (NaN/Inf in the input?)");
      return BK_ERR_NUMERIC;
    }
  tm.start();
  int nw = 0;
  StedcStats st;
  BK_TRY(stedc(ctx, n, dh.data(), eh.data(), ev.data(), max_want, rel_thresh, &nw, Z, ldz, &st));
  const double t_dc = tm.stop();
  if (Z && nw > n / 3 && n > kTwoStageFullMax && !getenv("BK_EIG_TWOSTAGE")) {
    *too_wide = true;
    return BK_OK;
  }
  tm.start();
  if (Z && nw > 0) BK_TRY(twostage_back(ctx, &ts, Z, ldz, nw));
  const double t_bt = tm.stop();
  if (getenv("BK_EIG_VERBOSE"))
    fprintf(stderr, "[eigen2 n=%d] sy2sb %.4f sb2st %.4f | dc %.4f | q2 %.4f q1 %.4f (k=%d)\n", n, ts.t_sy2sb,
            ts.t_sb2st, t_dc, ts.t_q2, ts.t_q1, nw);
  for (int i = 0; i < n; ++i) evals_host[i] = ev[n - 1 - i];
  if (n_want) *n_want = nw;
  if (times) {
    times->tridiag = t_tri;
    times->dc = t_dc;
    times->backtransform = t_bt;
    times->dc_stats = st;
    times->sytrd = SytrdStats();
    times->twostage = 1;
    times->t_sy2sb = ts.t_sy2sb;
    times->t_sb2st = ts.t_sb2st;
    times->t_q2 = ts.t_q2;
    times->t_q1 = ts.t_q1;
    times->band = ts.band;
  }
  return BK_OK;
}

// Distributed variant.  Stage 1 (dense -> band) is spread over the ranks; every rank then holds the complete band and
// runs the band -> tridiagonal stage and the divide & conquer itself (deterministic: the same bits everywhere, and
// cheaper than waiting for rank 0 and a broadcast of the stage-2 reflectors); the back-transformation - the part
// whose cost grows with the number of eigenvectors - is split by columns and the blocks are all-gathered through the
// symmetric buffer at heap offset off_Q (n x max_want doubles, ld n = ldz).  Every rank returns the same status.
int eigen_full_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n,
                    double* evals_host, int max_want, double rel_thresh, int* n_want, double* Z, long long ldz,
                    size_t off_Q, EigenTimes* times) {
  BK_REQUIRE(ldz == n, "eigen_full_dist: the eigenvector block must be dense (ld = n)");
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  DevBuf<double> d, e, agree;
  BK_TRY(d.alloc(n));
  BK_TRY(e.alloc(n));
  BK_TRY(agree.alloc(2));
  BK_CUDA(cudaMemsetAsync(e.p, 0, sizeof(double) * n, ctx->stream));
  TwoStage ts;
  tm.start();
  BK_TRY(twostage_reduce_dist(ctx, peer, X, ldx, p, sigma, n, &ts, d.p, e.p));
  const double t_tri = tm.stop();
  if (n_want) *n_want = 0;
  std::vector<double> dh(n), eh(n), ev(n);
  BK_CUDA(cudaMemcpyAsync(dh.data(), d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaMemcpyAsync(eh.data(), e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  int rc = BK_OK;
  for (int i = 0; i < n && rc == BK_OK; ++i)
    if (!std::isfinite(dh[i]) || !std::isfinite(eh[i])) {
      set_error("eigen: tridiagonalisation produced a non-finite entry (NaN/Inf in the input?)");
      rc = BK_ERR_NUMERIC;
    }
  tm.start();
  int nw = 0;
  StedcStats st;
  if (rc == BK_OK) {
    rc = stedc(ctx, n, dh.data(), eh.data(), ev.data(), max_want, rel_thresh, &nw, Z, ldz, &st);
    if (rc == BK_ERR_CUDA) return rc;
  }
  const double t_dc = tm.stop();
  // agreement before anybody enters the all-gather: a numerical failure (or a different count of retained
  // eigenvectors) on one rank must surface as the same error everywhere, not as a rank waiting in a collective
  {
    const double mine[2] = {rc != BK_OK ? 1.0 : 0.0, (double)nw};
    double all[2];
    BK_CUDA(cudaMemcpyAsync(agree.p, mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->stream));
    BK_TRY(peer_allreduce_sum(peer, agree.p, 2, ctx->stream));
    BK_CUDA(cudaMemcpyAsync(all, agree.p, sizeof(all), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    if (all[0] > 0.0) {
      if (rc == BK_OK) set_error("the eigensolver failed on another rank");
      return BK_ERR_NUMERIC;
    }
    if (all[1] != (double)nw * peer->world) {
      set_error("eigen_full_dist: the ranks disagree on the number of retained eigenvectors");
      return BK_ERR_NUMERIC;
    }
  }
  tm.start();
  if (Z && nw > 0) {
    const int G = peer->world, g = peer->rank;
    std::vector<long long> counts(G), displs(G);
    for (int r = 0; r < G; ++r) {
      const long long a = (long long)nw * r / G, b = (long long)nw * (r + 1) / G;
      counts[r] = (b - a) * n;
      displs[r] = a * n;
    }
    const int c0 = (int)((long long)nw * g / G), c1 = (int)((long long)nw * (g + 1) / G);
    double* Qs = peer_ptr(peer, off_Q);
    if (c1 > c0) {
      BK_TRY(twostage_back(ctx, &ts, Z + (size_t)c0 * ldz, ldz, c1 - c0));
      BK_CUDA(cudaMemcpyAsync(Qs + (size_t)c0 * n, Z + (size_t)c0 * ldz, sizeof(double) * (size_t)n * (c1 - c0),
                              cudaMemcpyDeviceToDevice, ctx->stream));
    }
    BK_TRY(peer_allgatherv_sym(peer, off_Q, counts.data(), displs.data(), ctx->stream));
    BK_CUDA(cudaMemcpyAsync(Z, Qs, sizeof(double) * (size_t)n * nw, cudaMemcpyDeviceToDevice, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  const double t_bt = tm.stop();
  for (int i = 0; i < n; ++i) evals_host[i] = ev[n - 1 - i];
  if (n_want) *n_want = nw;
  if (times) {
    times->tridiag = t_tri;
    times->twostage = 1;
    times->t_sy2sb = ts.t_sy2sb;
    times->band = ts.band;
    times->dc = t_dc;
    times->backtransform = t_bt;
    times->dc_stats = st;
    times->sytrd = SytrdStats();
    times->t_sb2st = ts.t_sb2st;
    times->t_q2 = ts.t_q2;
    times->t_q1 = ts.t_q1;
  }
  return BK_OK;
}

int eigen_full(bk_ctx* ctx, const double* K, long long ldk, int n, double* evals_host, int max_want,
               double rel_thresh, int* n_want, double* Z, long long ldz, EigenTimes* times) {
  Timer tm;
  BK_TRY(tm.init(ctx->stream));
  if (use_twostage(n, Z ? max_want : 0, rel_thresh)) {
    bool too_wide = false;
    BK_TRY(eigen_full_twostage(ctx, K, ldk, n, evals_host, max_want, rel_thresh, n_want, Z, ldz, times, &too_wide));
    if (!too_wide) return BK_OK;
  }
  DevBuf<double> d, e, tau, workbuf;
  const long long ldw = sytrd_ld(n);
  BK_TRY(workbuf.borrow(ctx->ws[0], (size_t)ldw * n));
  double* work = workbuf.p;
  BK_CUDA(cudaMemsetAsync(work, 0, sizeof(double) * (size_t)ldw * n, ctx->stream));
  BK_TRY(d.alloc(n));
  BK_TRY(e.alloc(n));
  BK_TRY(tau.alloc(n));
  BK_CUDA(cudaMemsetAsync(e.p, 0, sizeof(double) * n, ctx->stream));
  BK_CUDA(cudaMemsetAsync(tau.p, 0, sizeof(double) * n, ctx->stream));
  tm.start();
  BK_TRY(copy_matrix(ctx, K, ldk, n, n, 1.0, work, ldw));
  SytrdStats sst;
  BK_TRY(sytrd_lower(ctx, work, ldw, n, d.p, e.p, tau.p, 64, &sst));
  const double t_tri = tm.stop();
  if (getenv("BK_EIG_VERBOSE"))
    fprintf(stderr, "[eigen n=%d] tridiag %.4f s: panel kernels %.4f, trailing GEMMs %.4f, other %.4f\n", n, t_tri,
            sst.kernel_seconds, sst.update_seconds, t_tri - sst.kernel_seconds - sst.update_seconds);
  std::vector<double> dh(n), eh(n), ev(n);
  BK_CUDA(cudaMemcpyAsync(dh.data(), d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaMemcpyAsync(eh.data(), e.p, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i)
    if (!std::isfinite(dh[i]) || !std::isfinite(eh[i])) {
      set_error("eigen: tridiagonalisation produced a non-finite entry (NaN/Inf in the input?)");
      return BK_ERR_NUMERIC;
    }
  tm.start();
  int nw = 0;
  StedcStats st;
  BK_TRY(stedc(ctx, n, dh.data(), eh.data(), ev.data(), max_want, rel_thresh, &nw, Z, ldz, &st));
  const double t_dc = tm.stop();
  tm.start();
  if (Z && nw > 0) BK_TRY(ormtr_lower(ctx, work, ldw, n, tau.p, Z, ldz, nw));
  const double t_bt = tm.stop();
  for (int i = 0; i < n; ++i) evals_host[i] = ev[n - 1 - i];
  if (n_want) *n_want = nw;
  if (times) {
    times->tridiag = t_tri;
    times->dc = t_dc;
    times->backtransform = t_bt;
    times->dc_stats = st;
    times->sytrd = sst;
  }
  return BK_OK;
}

}  // namespace bk
