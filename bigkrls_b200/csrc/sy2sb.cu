// Two-stage tridiagonalisation, stage 1: dense symmetric -> band (bandwidth b = 64).
// Executable specification: tests/twostage_prototype.py (sy2sb).  Together with sb2st.cu this replaces the
// one-stage reduction of sytrd.cu when only part of the eigenvectors is wanted: all O(N^3) work becomes
// DMMA GEMMs (the one-stage algorithm is bound by an HBM-resident symmetric mat-vec per column).
//
// Per panel (columns c0 .. c0+b-1, rows r0 = c0+b .. n-1, m = n - r0):
//   1. Householder QR of the m x b panel in ONE cooperative kernel: the panel is distributed by row
//      slabs over the CTAs and stays in SHARED MEMORY for all b columns; per column one grid barrier
//      (partial dot products of the pivot column with the remaining columns -> every CTA forms tau and
//      the update coefficients redundantly -> updates its own slab).
//   2. T (compact WY) from V'V;  Z = A22 (V T);  W = Z - 1/2 V (T' (V'Z));  A22 -= [V W][W V]'
//      on the DMMA GEMM - 4 m^2 b useful flops per panel, (4/3) N^3 in total.  Panels are taken in PAIRS
//      (see the loop in sy2sb): only the second panel's columns get the first panel's update at once, the
//      second Z is corrected for the pending update, and one rank-4b update serves both panels.
// The matrix is kept fully symmetric (lower tiles are computed, the GEMM epilogue writes the transposed tile
// as well) so that A22 (V T) is a plain GEMM.
// V is left LAPACK-style below the band of A (unit diagonal implicit) and T_k is stored for the
// back-transformation.
#include <algorithm>
#include <cstdlib>
#include <utility>
#include <vector>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"
#include "peer.cuh"

namespace bk {

static constexpr int SB = 64;     // bandwidth / panel width
static constexpr int QR_NT = 256;
// CTAs of the cooperative panel factorisation per SM (tuning knob BK_QR_PER_SM, with BK_QR_ROWS = rows per CTA)
static int qr_ctas_per_sm() {
  static const int v = getenv("BK_QR_PER_SM") ? std::max(1, std::min(2, atoi(getenv("BK_QR_PER_SM")))) : 1;
  return v;
}

struct PanelArgs {
  double* A;        // n x n, lda
  long long lda;
  int n, c0;
  double* V;        // explicit V, rows r0.., ld = ldv (two copies: V and V2)
  double* V2;
  long long ldv;
  double* taus;     // SB
  double* part;     // [2][G][SB]
  double* prow;     // [2][SB]
  unsigned* barrier;
  long long* prof;  // optional (BK_QR_PROF), CTA 0: cycles in [0] partial dots, [1] barrier, [2] reduction, [3] update, [4] columns
  int rows_per;
};

__global__ void __launch_bounds__(QR_NT, 1) panel_qr_kernel(PanelArgs a) {
  extern __shared__ __align__(16) double slab[];  // rows_per x (SB+1)
  __shared__ double s_dots[SB], s_prow[SB];
  constexpr int LD = SB + 1;
  const int n = a.n, r0 = a.c0 + SB, m = n - r0;
  const int G = gridDim.x;
  const int row_lo = blockIdx.x * a.rows_per;              // local (panel) row range of this CTA
  const int nrows = max(0, min(a.rows_per, m - row_lo));
  const int nr = min(SB, m - 1);                           // reflectors
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned epoch = 0;

  for (int idx = threadIdx.x; idx < nrows * SB; idx += QR_NT) {
    const int r = idx % nrows, l = idx / nrows;
    slab[r * LD + l] = a.A[(long long)(r0 + row_lo + r) + (long long)(a.c0 + l) * a.lda];
  }
  __syncthreads();

  for (int j = 0; j < nr; ++j) {
    long long tc0 = 0, tc1 = 0, tc2 = 0, tc3 = 0;
    if (a.prof) tc0 = clock64();
    const int buf = j & 1;
    double* part = a.part + (size_t)buf * G * SB;
    double* prow = a.prow + buf * SB;
    // ---- partial sums over own rows strictly below the pivot row j -----------------------------------
    const int rs = max(0, j + 1 - row_lo);  // first own local row with panel row > j
    {
      // warp w owns columns l = j + w + 8 q (q < 8): one pass over the rows accumulates all of them, then the
      // eight warp reductions are interleaved (their shuffle latencies overlap)
      double acc[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) acc[q] = 0.0;
      for (int r = rs + lane; r < nrows; r += 32) {
        const double xj = slab[r * LD + j];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int l = j + warp + 8 * q;
          if (l < SB) acc[q] = fma(xj, slab[r * LD + l], acc[q]);
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
      }
      if (lane == 0) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int l = j + warp + 8 * q;
          if (l < SB) part[(size_t)blockIdx.x * SB + l] = acc[q];
        }
      }
    }
    if (j >= row_lo && j < row_lo + nrows) {
      for (int l = j + threadIdx.x; l < SB; l += QR_NT) prow[l] = slab[(j - row_lo) * LD + l];
    }
    if (a.prof) {
      __syncthreads();
      tc1 = clock64();
    }
    grid_barrier(a.barrier, epoch);
    if (a.prof) tc2 = clock64();
    // ---- every CTA: totals, reflector, update coefficients ---------------------------------------------
    {
      // each CTA keeps its partials in one contiguous 512-byte row (a transposed layout with coalesced
      // reads was measured slower: 16 CTAs then share every line they write)
      const int l = j + (threadIdx.x >> 2), sub = threadIdx.x & 3;
      double s = 0.0;
      if (l < SB) {
        // all loads of a thread in flight at once (40 covers 160 CTAs per round), summed in a fixed order
        for (int c0 = sub; c0 < G; c0 += 160) {
          double v[40];
#pragma unroll
          for (int q = 0; q < 40; ++q) {
            const int c = c0 + 4 * q;
            v[q] = (c < G) ? __ldcg(part + (size_t)c * SB + l) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 40; q += 8)
            s += ((v[q] + v[q + 1]) + (v[q + 2] + v[q + 3])) + ((v[q + 4] + v[q + 5]) + (v[q + 6] + v[q + 7]));
        }
      }
      s += __shfl_xor_sync(0xffffffffu, s, 1);
      s += __shfl_xor_sync(0xffffffffu, s, 2);
      if (l < SB && sub == 0) {
        s_dots[l] = s;
        s_prow[l] = __ldcg(prow + l);
      }
    }
    __syncthreads();
    if (a.prof) tc3 = clock64();
    const double alpha = s_prow[j], xnorm2 = s_dots[j];
    double beta, tau, scale;
    if (xnorm2 == 0.0) {
      beta = alpha;
      tau = 0.0;
      scale = 0.0;
    } else {
      beta = -copysign(sqrt(alpha * alpha + xnorm2), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) a.taus[j] = tau;
    // own rows below the pivot: P[r,l] -= v_r * t_l, then column j becomes v
    {
      // thread = (column l, row phase): no integer division in the loop, consecutive lanes on consecutive columns
      const int l = j + 1 + (threadIdx.x & (SB - 1)), rph = threadIdx.x >> 6;
      if (l < SB) {
        const double tl = tau * (s_prow[l] + scale * s_dots[l]);
        const double stl = scale * tl;
#pragma unroll 4
        for (int r = rs + rph; r < nrows; r += QR_NT / SB) slab[r * LD + l] = fma(-slab[r * LD + j], stl, slab[r * LD + l]);
      }
    }
    __syncthreads();
    for (int r = rs + threadIdx.x; r < nrows; r += QR_NT) slab[r * LD + j] *= scale;
    if (j >= row_lo && j < row_lo + nrows) {  // pivot row: v = 1
      for (int l = j + 1 + threadIdx.x; l < SB; l += QR_NT)
        slab[(j - row_lo) * LD + l] -= tau * (s_prow[l] + scale * s_dots[l]);
      if (threadIdx.x == 0) slab[(j - row_lo) * LD + j] = beta;
    }
    __syncthreads();
    if (a.prof && blockIdx.x == 0 && threadIdx.x == 0) {
      const long long tc4 = clock64();
      a.prof[0] += tc1 - tc0;
      a.prof[1] += tc2 - tc1;
      a.prof[2] += tc3 - tc2;
      a.prof[3] += tc4 - tc3;
      a.prof[4] += 1;
    }
  }
  if (blockIdx.x == 0)
    for (int l = nr + threadIdx.x; l < SB; l += QR_NT) a.taus[l] = 0.0;
  // ---- write back: LAPACK-style into A, explicit unit-lower-trapezoidal V (two copies) ----------------
  for (int idx = threadIdx.x; idx < nrows * SB; idx += QR_NT) {
    const int r = idx % nrows, l = idx / nrows;
    const int pr = row_lo + r;  // panel row
    const double x = slab[r * LD + l];
    a.A[(long long)(r0 + pr) + (long long)(a.c0 + l) * a.lda] = x;
    double v = 0.0;
    if (l < nr) v = (pr > l) ? x : (pr == l ? 1.0 : 0.0);
    a.V[(long long)(r0 + pr) + (long long)l * a.ldv] = v;
    a.V2[(long long)(r0 + pr) + (long long)l * a.ldv] = v;
  }
}

// T (ib x ib upper triangular) from S = V'V and tau: T[:i,i] = -tau_i T[:i,:i] S[:i,i].  16 threads per row of T
// share each dot product (the 64 steps are sequential; a single thread per row made a step cost a 64-long
// dependent FMA chain).  Launch with 16 * ib threads.
__global__ void sb_larft_kernel(const double* __restrict__ S, const double* __restrict__ tau, int ib,
                                double* __restrict__ T) {
  extern __shared__ double ts[];  // T (ib x ib) then S (ib x ib)
  double* ss = ts + ib * ib;
  const int r = threadIdx.x >> 4, part = threadIdx.x & 15;
  for (int idx = threadIdx.x; idx < ib * ib; idx += blockDim.x) {
    ts[idx] = 0.0;
    ss[idx] = S[idx];
  }
  __syncthreads();
  __shared__ double s_tau[64];
  if (threadIdx.x < ib) s_tau[threadIdx.x] = tau[threadIdx.x];
  __syncthreads();
  for (int i = 0; i < ib; ++i) {
    const double ti = s_tau[i];
    double acc = 0.0;
    if (r < i)
      for (int q = r + part; q < i; q += 16) acc = fma(ts[r * ib + q], ss[q + i * ib], acc);  // ts holds T row-major
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    __syncthreads();
    if (part == 0) {
      if (r < i) ts[r * ib + i] = -ti * acc;
      if (r == i) ts[i * ib + i] = ti;
    }
    __syncthreads();
  }
  for (int idx = threadIdx.x; idx < ib * ib; idx += blockDim.x) T[idx] = ts[(idx % ib) * ib + idx / ib];
}

// The same T by inversion: for H = H_1 ... H_b = I - V T V' one has T^-1 = striu(V'V) + diag(1 / tau) exactly (for any
// tau, orthogonal or not), so T is the inverse of an upper triangular matrix - built by doubling: with the inverses of
// two neighbouring s x s diagonal blocks (A^-1, C^-1) and the block B above-right of them, the inverse of the 2s x 2s
// block has -A^-1 B C^-1 there.  Six levels of small products with two barriers each instead of 64 dependent steps
// (40 us -> ~8 us per panel, and T sits on the critical path of every panel).  A reflector with tau = 0 (H_i = I) is
// decoupled (its row and column of T are zero).  One CTA of 256 threads, ib <= 64.
__global__ void __launch_bounds__(256) sb_larft_inv_kernel(const double* __restrict__ S, const double* __restrict__ tau,
                                                           int ib, double* __restrict__ T) {
  extern __shared__ __align__(16) double li_sm[];
  double (*U)[SB + 1] = reinterpret_cast<double (*)[SB + 1]>(li_sm);
  double (*R)[SB + 1] = U + SB;
  double (*Y)[SB + 1] = R + SB;
  __shared__ double s_tau[SB];
  const int tid = threadIdx.x;
  if (tid < SB) s_tau[tid] = (tid < ib) ? tau[tid] : 0.0;
  __syncthreads();
  for (int idx = tid; idx < SB * SB; idx += 256) {
    const int i = idx % SB, j = idx / SB;
    const bool live = s_tau[i] != 0.0 && s_tau[j] != 0.0;
    double u = 0.0;
    if (i == j)
      u = (s_tau[i] != 0.0) ? 1.0 / s_tau[i] : 1.0;
    else if (i < j && live)
      u = S[i + j * ib];
    U[i][j] = u;
    R[i][j] = (i == j) ? ((s_tau[i] != 0.0) ? s_tau[i] : 1.0) : 0.0;
  }
  __syncthreads();
  for (int sz = 1; sz < SB; sz *= 2) {
    // element e of this level: pair p = e / (sz*sz), (i, j) inside the off-diagonal block
    const int per = sz * sz, total = (SB / (2 * sz)) * per;
    for (int e = tid; e < total; e += 256) {
      const int pblk = e / per, i = (e % per) % sz, j = (e % per) / sz;
      const int b0 = pblk * 2 * sz;
      double acc = 0.0;  // Y = B C^-1 (C^-1 upper triangular: k <= j)
      for (int k = 0; k <= j; ++k) acc = fma(U[b0 + i][b0 + sz + k], R[b0 + sz + k][b0 + sz + j], acc);
      Y[b0 + i][b0 + sz + j] = acc;
    }
    __syncthreads();
    for (int e = tid; e < total; e += 256) {
      const int pblk = e / per, i = (e % per) % sz, j = (e % per) / sz;
      const int b0 = pblk * 2 * sz;
      double acc = 0.0;  // X = -A^-1 Y (A^-1 upper triangular: k >= i)
      for (int k = i; k < sz; ++k) acc = fma(R[b0 + i][b0 + k], Y[b0 + k][b0 + sz + j], acc);
      R[b0 + i][b0 + sz + j] = -acc;
    }
    __syncthreads();
  }
  for (int idx = tid; idx < ib * ib; idx += 256) {
    const int i = idx % ib, j = idx / ib;
    double t = 0.0;
    if (i == j)
      t = s_tau[i];
    else if (i < j && s_tau[i] != 0.0 && s_tau[j] != 0.0)
      t = R[i][j];
    T[idx] = t;
  }
}

// AB[d + j*ldab] = A[j+d, j] for d <= b, 0 for b < d < ldab   (lower band, working width 2b)
__global__ void extract_band_kernel(const double* __restrict__ A, long long lda, int n, int b,
                                    double* __restrict__ AB, int ldab) {
  const long long total = (long long)n * ldab;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(idx % ldab), j = (int)(idx / ldab);
    double v = 0.0;
    if (d <= b && j + d < n) v = A[(long long)(j + d) + (long long)j * lda];
    AB[idx] = v;
  }
}

// T from V'V and tau (one CTA): the doubling inverse by default, the 64-step recurrence with BK_LARFT_SEQ=1
static void launch_larft(cudaStream_t st, const double* S, const double* tau, int ib, double* T) {
  static const bool seq = getenv("BK_LARFT_SEQ") != nullptr;
  if (seq)
    sb_larft_kernel<<<1, 16 * ib, sizeof(double) * 2 * ib * ib, st>>>(S, tau, ib, T);
  else {
    constexpr size_t smem = sizeof(double) * 3 * SB * (SB + 1);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(sb_larft_inv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      attr_set = true;
    }
    sb_larft_inv_kernel<<<1, 256, smem, st>>>(S, tau, ib, T);
  }
}
// W = Z - V (T' V'Z) / 2 in place of Z (Wd, rows r0..) and a copy into Wd2.  (One fused kernel for the last three
// steps - every CTA forming T' V'Z itself, FMA loops - was measured slower than the three launches: +0.01 s per fit.)
static int finish_w_fused(bk_ctx* ctx, int n, int r0, const double* Vd, double* Wd, double* Wd2, const double* Tk,
                          double* S2, double* S3) {
  const int b = SB, m = n - r0;
  BK_TRY(gemm(ctx, true, false, b, b, m, 1.0, Vd + r0, n, Wd + r0, n, 0.0, S2, b));   // V'Z
  BK_TRY(gemm(ctx, true, false, b, b, b, 1.0, Tk, b, S2, b, 0.0, S3, b));             // T' V'Z
  BK_TRY(gemm(ctx, false, false, m, b, b, -0.5, Vd + r0, n, S3, b, 1.0, Wd + r0, n)); // W = Z - V S3 / 2
  BK_TRY(copy_matrix(ctx, Wd + r0, n, m, b, 1.0, Wd2 + r0, n));
  return BK_OK;
}

// Look-ahead variant: one panel at a time, and the factorisation of panel k+1 (cooperative QR kernel, V'V, T, V T -
// all latency-bound) runs on the high-priority side stream UNDER the trailing update of panel k:
//   main : Z_k = A22 (V T)_k -> W_k -> update of the NEXT panel's columns only -> [event] -> rest of the update
//   side :                                                         [wait] QR_{k+1}, T_{k+1}, (V T)_{k+1} [event]
// The update splits into the next panel's b columns (all rows) and the trailing (m-b) x (m-b) block (lower tiles,
// mirrored); the mirror image of the panel columns is never read again.  [V W] / [W V] are double-buffered by
// panel parity because QR_{k+1} writes V_{k+1} while the update still reads V_k, W_k.
static int sy2sb_lookahead(bk_ctx* ctx, double* A, long long lda, int n, double* Tstore, double* AB, int ldab,
                           BandStats* stats) {
  const int b = SB;
  const int G = ctx->sm_count;
  DevBuf<double> taus, part, prow, S, VT, S2, S3, PAbuf, PBbuf;
  BK_TRY(taus.alloc(b));
  BK_TRY(part.alloc((size_t)2 * G * qr_ctas_per_sm() * b));
  BK_TRY(prow.alloc(2 * b));
  BK_TRY(S.alloc(b * b));
  BK_TRY(VT.alloc((size_t)n * b));
  BK_TRY(S2.alloc(b * b));
  BK_TRY(S3.alloc(b * b));
  BK_TRY(PAbuf.borrow(ctx->panel_cache[0], (size_t)4 * b * n));
  BK_TRY(PBbuf.borrow(ctx->panel_cache[1], (size_t)4 * b * n));
  BK_TRY(ctx->barrier.ensure(4));
  static const int rows_target = getenv("BK_QR_ROWS") ? atoi(getenv("BK_QR_ROWS")) : 128;
  const int max_rows_per = std::max(rows_target, (int)ceil_div(std::max(1, n - b), (int64_t)G * qr_ctas_per_sm()));
  const size_t max_smem = (size_t)max_rows_per * (b + 1) * sizeof(double);
  BK_REQUIRE(max_smem <= 200 * 1024, "sy2sb: n too large for the shared-memory panel slabs");
  BK_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  BK_CUDA(cudaFuncSetAttribute(sb_larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(double) * 2 * b * b)));
  BK_CUDA(cudaMemsetAsync(PAbuf.p, 0, sizeof(double) * (size_t)4 * b * n, ctx->stream));
  BK_CUDA(cudaMemsetAsync(PBbuf.p, 0, sizeof(double) * (size_t)4 * b * n, ctx->stream));
  const size_t blk = (size_t)b * n;
  cudaStream_t main_st = ctx->stream, side_st = ctx->side_stream;
  // events: 0/1 ordering (columns ready, panel factored), 2.. timing pairs of the large GEMMs
  cudaEvent_t e_col = pool_event(ctx, 0), e_fact = pool_event(ctx, 1);
  size_t n_ev = 2;
  double flops = 0.0;
  auto mark = [&]() {
    if (!stats) return;
    cudaEventRecord(pool_event(ctx, n_ev++), ctx->stream);
  };

  // QR of the panel at columns c0 on the CURRENT ctx->stream: V -> Vd, Vd2; T -> Tk; VT = V T
  auto factor_panel = [&](int c0, double* Vd, double* Vd2, double* Tk) -> int {
    const int r0 = c0 + b, m = n - r0;
    PanelArgs pa;
    pa.A = A;
    pa.lda = lda;
    pa.n = n;
    pa.c0 = c0;
    pa.V = Vd;
    pa.V2 = Vd2;
    pa.ldv = n;
    pa.taus = taus.p;
    pa.part = part.p;
    pa.prow = prow.p;
    pa.barrier = ctx->barrier.p;
    pa.prof = nullptr;
    const int Gp = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)G * qr_ctas_per_sm(), ceil_div(m, rows_target)));
    pa.rows_per = (int)ceil_div(m, Gp);
    BK_CUDA(cudaMemsetAsync(ctx->barrier.p, 0, sizeof(unsigned) * 4, ctx->stream));
    void* kargs[] = {&pa};
    const size_t smem = (size_t)pa.rows_per * (b + 1) * sizeof(double);
    BK_CUDA(cudaLaunchCooperativeKernel((void*)panel_qr_kernel, dim3(Gp), dim3(QR_NT), kargs, smem, ctx->stream));
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, true, false, b, b, m, 1.0, Vd + r0, n, Vd + r0, n, 0.0, S.p, b));
    launch_larft(ctx->stream, S.p, taus.p, b, Tk);
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, false, false, m, b, b, 1.0, Vd + r0, n, Tk, b, 0.0, VT.p, m));
    return BK_OK;
  };

  int k = 0;
  bool on_side = false;  // the current panel was factored on the side stream
  if (n - b >= 2) BK_TRY(factor_panel(0, PAbuf.p, PBbuf.p + blk, Tstore));
  for (int c0 = 0; c0 < n; c0 += b, ++k) {
    const int r0 = c0 + b, m = n - r0;
    if (m < 2) break;
    const int q = k & 1;
    double* PA = PAbuf.p + (size_t)q * 2 * blk;  // [V | W] of this panel
    double* PB = PBbuf.p + (size_t)q * 2 * blk;  // [W | V]
    double* Tk = Tstore + (size_t)k * b * b;
    double* A22 = A + r0 + (long long)r0 * lda;
    if (on_side) BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
    mark();
    BK_TRY(gemm(ctx, false, false, m, b, m, 1.0, A22, lda, VT.p, m, 0.0, PA + blk + r0, n));  // Z = A22 V T
    mark();
    flops += 2.0 * m * (double)m * b;
    // W = Z - V (T' V'Z) / 2 in place, and the copy into [W | V]
    BK_TRY(finish_w_fused(ctx, n, r0, PA, PA + blk, PB, Tk, S2.p, S3.p));
    const int r1 = r0 + b, m2 = n - r1;
    if (m2 < 2) {
      // last panel: the whole trailing block (its lower triangle carries the final diagonal blocks of the band)
      mark();
      BK_TRY(gemm(ctx, false, true, m, m, 2 * b, -1.0, PA + r0, n, PB + r0, n, 1.0, A22, lda, 2));
      mark();
      flops += 1.0 * m * (double)m * 2 * b;
      on_side = false;
      continue;
    }
    // the next panel's columns (rows r0.., incl. its diagonal block) get the update first ...
    BK_TRY(gemm(ctx, false, true, m, b, 2 * b, -1.0, PA + r0, n, PB + r0, n, 1.0, A22, lda));
    BK_CUDA(cudaEventRecord(e_col, main_st));
    // ... so that its factorisation can start on the side stream while the rest of the update runs here
    {
      BK_CUDA(cudaStreamWaitEvent(side_st, e_col, 0));
      SideStreamScope sc(ctx);
      double* PAn = PAbuf.p + (size_t)(q ^ 1) * 2 * blk;
      double* PBn = PBbuf.p + (size_t)(q ^ 1) * 2 * blk;
      BK_TRY(factor_panel(r0, PAn, PBn + blk, Tstore + (size_t)(k + 1) * b * b));
      BK_CUDA(cudaEventRecord(e_fact, side_st));
    }
    on_side = true;
    mark();
    BK_TRY(gemm(ctx, false, true, m2, m2, 2 * b, -1.0, PA + r1, n, PB + r1, n, 1.0, A + r1 + (long long)r1 * lda, lda, 2));
    mark();
    flops += 1.0 * m2 * (double)m2 * 2 * b;
  }
  if (on_side) BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
  extract_band_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n * ldab, 256), 16LL * ctx->sm_count), 256, 0,
                        ctx->stream>>>(A, lda, n, b, AB, ldab);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (stats) {
    double sec = 0.0;
    for (size_t i = 2; i + 1 < n_ev; i += 2) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ctx->event_pool[i], ctx->event_pool[i + 1]);
      sec += ms * 1e-3;
    }
    stats->gemm_launches = (double)((n_ev - 2) / 2);
    stats->gemm_seconds = sec;
    stats->gemm_flops = flops;
  }
  return BK_OK;
}

// Ksrc (optional): the symmetric input matrix, read-only.  With it A need not hold a copy on entry, and for large n the
// first part of the reduction runs as a PIPELINE (see "delayed update" below); without it A holds the matrix and
// everything runs in place.
int sy2sb(bk_ctx* ctx, double* A, long long lda, int n, double* Tstore, double* AB, int ldab, BandStats* stats,
          const double* Ksrc, long long ldk) {
  const int b = SB;
  BK_REQUIRE(ldab >= 2 * b, "sy2sb: band storage needs 2b rows");
  if (const char* la = getenv("BK_SY2SB_LOOKAHEAD"))
    if (atoi(la) != 0) return sy2sb_lookahead(ctx, A, lda, n, Tstore, AB, ldab, stats);
  const int G = ctx->sm_count;
  DevBuf<double> taus, part, prow, S, T, VT, S2, S3;
  BK_TRY(taus.alloc(b));
  BK_TRY(part.alloc((size_t)2 * G * qr_ctas_per_sm() * b));
  BK_TRY(prow.alloc(2 * b));
  BK_TRY(S.alloc(b * b));
  BK_TRY(T.alloc(b * b));
  BK_TRY(VT.alloc((size_t)n * b));
  BK_TRY(S2.alloc(b * b));
  BK_TRY(S3.alloc(b * b));
  BK_TRY(ctx->barrier.ensure(4));
  // Rows of the panel per CTA (lower bound; all SMs are used while m >= rows_target * SMs).  Measured: the
  // per-column cost is dominated by the per-CTA slab work, not by the grid barrier - thin slabs win.
  static const int rows_target = getenv("BK_QR_ROWS") ? atoi(getenv("BK_QR_ROWS")) : 128;
  const int max_rows_per = std::max(rows_target, (int)ceil_div(std::max(1, n - b), (int64_t)G * qr_ctas_per_sm()));
  const size_t max_smem = (size_t)max_rows_per * (b + 1) * sizeof(double);
  BK_REQUIRE(max_smem <= 200 * 1024, "sy2sb: n too large for the shared-memory panel slabs");
  BK_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  BK_CUDA(cudaFuncSetAttribute(sb_larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(double) * 2 * b * b)));
  static const bool full_update = getenv("BK_SY2SB_FULL") != nullptr;
  // CUDA events around the two large GEMMs of every panel (roofline of the dominant kernel, bench.py)
  // (events come from the context's pool: created once, reused by every fit)
  size_t n_ev = 16;  // 0..15: ordering events of the pipelined phase; then start/end pairs of the large GEMMs
  const size_t ev0 = n_ev;
  double flops = 0.0;
  auto mark = [&]() {
    if (!stats) return;
    cudaEventRecord(pool_event(ctx, n_ev++), ctx->stream);
  };
  DevBuf<long long> prof;
  if (getenv("BK_QR_PROF")) {
    BK_TRY(prof.alloc(8));
    BK_CUDA(cudaMemsetAsync(prof.p, 0, 8 * sizeof(long long), ctx->stream));
  }
  // Panels are processed in PAIRS so that the big symmetric update runs with k = 4b = 256 instead of 2b (28 TF
  // against 22.8 TF measured): after the first panel only the columns of the second panel are updated; the
  // product Z2 = A22 (V2 T2) of the second panel is corrected for the pending update
  //   Z2 -= [V1 W1] ([W1 V1]' V2 T2),
  // and one rank-4b update A22 -= [V1 W1 V2 W2][W1 V1 W2 V2]' follows.  PA = [V1|W1|V2|W2], PB = [W1|V1|W2|V2].
  static const bool pair_panels = getenv("BK_SY2SB_NOPAIR") == nullptr;
  // Look-ahead: the FIRST panel of the next pair is factored on the high-priority side stream under the rank-4b
  // update of this pair (its columns get the update first).  [V W] / [W V] are therefore double-buffered by pair
  // parity: QR writes the next V1 while the update still reads this pair's panels.
  static const bool pair_lookahead = getenv("BK_SY2SB_NOLA") == nullptr;
  DevBuf<double> PAbuf, PBbuf, Gs;
  BK_TRY(PAbuf.borrow(ctx->panel_cache[0], (size_t)8 * b * n));
  BK_TRY(PBbuf.borrow(ctx->panel_cache[1], (size_t)8 * b * n));
  BK_TRY(Gs.alloc((size_t)2 * b * b));
  BK_CUDA(cudaMemsetAsync(PAbuf.p, 0, sizeof(double) * (size_t)8 * b * n, ctx->stream));
  BK_CUDA(cudaMemsetAsync(PBbuf.p, 0, sizeof(double) * (size_t)8 * b * n, ctx->stream));
  const size_t blk = (size_t)b * n;
  cudaStream_t main_st = ctx->stream, side_st = ctx->side_stream;
  cudaEvent_t e_col = nullptr, e_fact = nullptr;
  BK_CUDA(cudaEventCreateWithFlags(&e_col, cudaEventDisableTiming));
  BK_CUDA(cudaEventCreateWithFlags(&e_fact, cudaEventDisableTiming));
  struct EvGuard {
    cudaEvent_t a, b;
    ~EvGuard() {
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  } evguard{e_col, e_fact};

  // Householder QR of the panel at columns c0 (rows r0 = c0 + b ..): V into Vd and Vd2, T into Tk, VT = V T
  auto factor_panel = [&](int c0, double* Vd, double* Vd2, double* Tk) -> int {
    const int r0 = c0 + b, m = n - r0;
    PanelArgs pa;
    pa.A = A;
    pa.lda = lda;
    pa.n = n;
    pa.c0 = c0;
    pa.V = Vd;
    pa.V2 = Vd2;
    pa.ldv = n;
    pa.taus = taus.p;
    pa.part = part.p;
    pa.prow = prow.p;
    pa.barrier = ctx->barrier.p;
    pa.prof = prof.p;
    const int Gp = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)G * qr_ctas_per_sm(), ceil_div(m, rows_target)));
    pa.rows_per = (int)ceil_div(m, Gp);
    BK_CUDA(cudaMemsetAsync(ctx->barrier.p, 0, sizeof(unsigned) * 4, ctx->stream));
    void* kargs[] = {&pa};
    const size_t smem = (size_t)pa.rows_per * (b + 1) * sizeof(double);
    BK_CUDA(cudaLaunchCooperativeKernel((void*)panel_qr_kernel, dim3(Gp), dim3(QR_NT), kargs, smem, ctx->stream));
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, true, false, b, b, m, 1.0, Vd + r0, n, Vd + r0, n, 0.0, S.p, b));
    launch_larft(ctx->stream, S.p, taus.p, b, Tk);
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, false, false, m, b, b, 1.0, Vd + r0, n, Tk, b, 0.0, VT.p, m));  // V T
    return BK_OK;
  };
  // W = Z - V (T' V'Z) / 2 in place of Z (Wd), and a copy into Wd2
  auto finish_w = [&](int r0, const double* Vd, double* Wd, double* Wd2, const double* Tk) -> int {
    return finish_w_fused(ctx, n, r0, Vd, Wd, Wd2, Tk, S2.p, S3.p);
  };

  int k = 0, pair = 0, c_start = 0;
  // ---- Delayed update (pipelined phase, large trailing matrices) -------------------------------------------------
  // In place, the two panel factorisations of a pair, their small GEMMs and the big products Z = A22 (V T) all wait for
  // the rank-256 update of the previous pair, and the update waits for them: per pair ~0.75 ms of latency-bound work
  // with the tensor pipe idle.  Here the chain of pair i does not wait for the update U_{i-1}: it works on the matrix as
  // it was BEFORE that update, A^(i-1), and corrects for it with skinny products:
  //     columns of the pair:  A^(i)[:, cols] = A^(i-1)[:, cols] - PA_{i-1} PB_{i-1}[cols]'             (into A)
  //     Z = A^(i) X           = A^(i-1) X - PA_{i-1} (PB_{i-1}' X)
  // while U_{i-1}: A^(i) = A^(i-1) - PA_{i-1} PB_{i-1}' runs on the other stream OUT OF PLACE into the second of two
  // work matrices (nobody overwrites what the chain still reads).  The chain runs on the high-priority stream; the
  // update fills the machine whenever the chain leaves it idle.  A receives only the factored panels (what the band
  // extraction and the back-transformation read); the last update of the phase writes the trailing matrix into A and
  // the in-place loop below takes over (small matrices are latency-bound either way, and there both panels of a pair
  // would be exposed).
  // (read per call: the tests run both variants in one process)
  const int pipe_min = getenv("BK_SY2SB_PIPE_MIN") ? atoi(getenv("BK_SY2SB_PIPE_MIN")) : 10240;
  const bool no_pipe = getenv("BK_SY2SB_NOPIPE") != nullptr;
  bool piped = Ksrc && pair_panels && !full_update && !no_pipe && (n - 3 * b >= pipe_min) && pipe_min >= 4 * b;
  if (piped) {
    DevBuf<double> W0, W1, Gs4;
    BK_TRY(W0.borrow(ctx->ws[2], (size_t)n * n));
    BK_TRY(W1.borrow(ctx->ws[3], (size_t)n * n));
    BK_TRY(Gs4.alloc((size_t)4 * b * b));
    double* Wb[2] = {W0.p, W1.p};
    // A^(m) lives in: the input for m = 0, else work matrix (m-1) & 1
    auto base_of = [&](int mth, long long& ld) -> const double* {
      if (mth == 0) {
        ld = ldk;
        return Ksrc;
      }
      ld = n;
      return Wb[(mth - 1) & 1];
    };
    auto e_chain = [&](int i) { return pool_event(ctx, 2 + (i & 3)); };
    auto e_upd = [&](int i) { return pool_event(ctx, 6 + (i & 3)); };
    cudaEvent_t e_begin = pool_event(ctx, 10);
    BK_CUDA(cudaEventRecord(e_begin, main_st));
    BK_CUDA(cudaStreamWaitEvent(side_st, e_begin, 0));
    // BK_SY2SB_PIPE_TRACE=i0: start/end of the chain and of the update of pairs i0..i0+7, relative to the phase's start
    const int tr_i0 = getenv("BK_SY2SB_PIPE_TRACE") ? atoi(getenv("BK_SY2SB_PIPE_TRACE")) : -1;
    std::vector<cudaEvent_t> trev;
    if (tr_i0 >= 0) {
      trev.resize(8 * 6);
      for (auto& e : trev) cudaEventCreate(&e);
    }
    auto trace_ev = [&](int i, int which, cudaStream_t st) {
      if (tr_i0 >= 0 && i >= tr_i0 && i < tr_i0 + 8) cudaEventRecord(trev[(size_t)(i - tr_i0) * 6 + which], st);
    };
    int i = 0;
    for (;; ++i) {
      const int c0 = 2 * b * i, r0 = c0 + b, r1 = r0 + b, r2 = r1 + b;
      const int m = n - r0, m2 = n - r1, m3 = n - r2;
      double* PA = PAbuf.p + (size_t)(i & 1) * 4 * blk;
      double* PB = PBbuf.p + (size_t)(i & 1) * 4 * blk;
      const double* PAp = PAbuf.p + (size_t)((i + 1) & 1) * 4 * blk;  // the pending pair i-1
      const double* PBp = PBbuf.p + (size_t)((i + 1) & 1) * 4 * blk;
      double* T1 = Tstore + (size_t)(2 * i) * b * b;
      double* T2 = T1 + (size_t)b * b;
      long long ldp = 0, ldi = 0;
      const double* Bp = base_of(i > 0 ? i - 1 : 0, ldp);  // A^(i-1) (i = 0: the input itself, nothing pending)
      const bool last = (m3 - 2 * b < pipe_min);           // the next pair is left to the in-place loop
      {
        SideStreamScope sc(ctx);
        if (i >= 2) BK_CUDA(cudaStreamWaitEvent(side_st, e_upd(i - 2), 0));  // A^(i-1) is complete
        trace_ev(i, 0, side_st);
        // columns of both panels (rows c0.., diagonal blocks included) into A
        if (i == 0)
          BK_TRY(copy_matrix(ctx, Ksrc, ldk, n, 2 * b, 1.0, A, lda));
        else
          BK_TRY(gemm_oop(ctx, false, true, n - c0, 2 * b, 4 * b, -1.0, PAp + c0, n, PBp + c0, n, 1.0,
                          Bp + c0 + (long long)c0 * ldp, ldp, A + c0 + (long long)c0 * lda, lda));
        // ---- first panel
        BK_TRY(factor_panel(c0, PA, PB + blk, T1));
        mark();
        BK_TRY(gemm(ctx, false, false, m, b, m, 1.0, Bp + r0 + (long long)r0 * ldp, ldp, VT.p, m, 0.0, PA + blk + r0, n));
        mark();
        flops += 2.0 * m * (double)m * b;
        if (i > 0) {
          BK_TRY(gemm(ctx, true, false, 4 * b, b, m, 1.0, PBp + r0, n, VT.p, m, 0.0, Gs4.p, 4 * b));
          BK_TRY(gemm(ctx, false, false, m, b, 4 * b, -1.0, PAp + r0, n, Gs4.p, 4 * b, 1.0, PA + blk + r0, n));
        }
        BK_TRY(finish_w(r0, PA, PA + blk, PB, T1));
        trace_ev(i, 1, side_st);
        // ---- second panel: its columns get the first panel's update, then as above with one more correction
        BK_TRY(gemm(ctx, false, true, m, b, 2 * b, -1.0, PA + r0, n, PB + r0, n, 1.0, A + r0 + (long long)r0 * lda, lda));
        BK_TRY(factor_panel(r0, PA + 2 * blk, PB + 3 * blk, T2));
        mark();
        BK_TRY(gemm(ctx, false, false, m2, b, m2, 1.0, Bp + r1 + (long long)r1 * ldp, ldp, VT.p, m2, 0.0, PA + 3 * blk + r1, n));
        mark();
        flops += 2.0 * m2 * (double)m2 * b;
        if (i > 0) {
          BK_TRY(gemm(ctx, true, false, 4 * b, b, m2, 1.0, PBp + r1, n, VT.p, m2, 0.0, Gs4.p, 4 * b));
          BK_TRY(gemm(ctx, false, false, m2, b, 4 * b, -1.0, PAp + r1, n, Gs4.p, 4 * b, 1.0, PA + 3 * blk + r1, n));
        }
        BK_TRY(gemm(ctx, true, false, 2 * b, b, m2, 1.0, PB + r1, n, VT.p, m2, 0.0, Gs.p, 2 * b));
        BK_TRY(gemm(ctx, false, false, m2, b, 2 * b, -1.0, PA + r1, n, Gs.p, 2 * b, 1.0, PA + 3 * blk + r1, n));
        BK_TRY(finish_w(r1, PA + 2 * blk, PA + 3 * blk, PB + 2 * blk, T2));
        trace_ev(i, 2, side_st);
        BK_CUDA(cudaEventRecord(e_chain(i), side_st));
      }
      // ---- U_i on the main stream: A^(i+1) = A^(i) - PA PB' on the trailing block, out of place
      BK_CUDA(cudaStreamWaitEvent(main_st, e_chain(i), 0));
      const double* Bi = base_of(i, ldi);
      double* Cout = last ? A + r2 + (long long)r2 * lda : Wb[i & 1] + r2 + (long long)r2 * n;
      const long long ldo = last ? lda : (long long)n;
      trace_ev(i, 3, main_st);
      mark();
      BK_TRY(gemm_oop(ctx, false, true, m3, m3, 4 * b, -1.0, PA + r2, n, PB + r2, n, 1.0, Bi + r2 + (long long)r2 * ldi, ldi,
                      Cout, ldo, 2));
      mark();
      flops += 1.0 * m3 * (double)m3 * 4 * b;
      trace_ev(i, 4, main_st);
      if (last) {
        // the first panel's columns of the next pair (rows r1.., its diagonal block included) with this pair's update
        BK_TRY(gemm_oop(ctx, false, true, m2, b, 4 * b, -1.0, PA + r1, n, PB + r1, n, 1.0, Bi + r1 + (long long)r1 * ldi, ldi,
                        A + r1 + (long long)r1 * lda, lda));
        ++i;
        break;
      }
      BK_CUDA(cudaEventRecord(e_upd(i), main_st));
    }
    if (tr_i0 >= 0) {
      cudaStreamSynchronize(main_st);
      cudaStreamSynchronize(side_st);
      for (int q = 0; q < 8 && tr_i0 + q < i; ++q) {
        float t[5];
        for (int w = 0; w < 5; ++w) cudaEventElapsedTime(&t[w], e_begin, trev[(size_t)q * 6 + w]);
        fprintf(stderr, "[sy2sb pipe] pair %d (m = %d): chain %.3f .. panel 2 at %.3f .. %.3f ms | update %.3f .. %.3f ms\n", tr_i0 + q,
                n - (2 * b * (tr_i0 + q) + b), t[0], t[1], t[2], t[3], t[4]);
      }
      for (auto& e : trev) cudaEventDestroy(e);
    }
    c_start = 2 * b * i;
    k = 2 * i;
    pair = i;
  } else if (Ksrc) {
    BK_TRY(copy_matrix(ctx, Ksrc, ldk, n, n, 1.0, A, lda));
  }
  bool ahead = false;  // the first panel of this pair was factored on the side stream during the previous update
  for (int c0 = c_start; c0 < n; ++pair) {
    const int r0 = c0 + b, m = n - r0;
    if (m < 2) break;
    double* PA = PAbuf.p + (size_t)(pair & 1) * 4 * blk;
    double* PB = PBbuf.p + (size_t)(pair & 1) * 4 * blk;
    double* A22 = A + r0 + (long long)r0 * lda;
    double* T1 = Tstore + (size_t)k * b * b;
    // ---- first panel of the pair: V1 -> PA[0], PB[1]; W1 -> PA[1], PB[0]
    if (ahead)
      BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
    else
      BK_TRY(factor_panel(c0, PA, PB + blk, T1));
    ahead = false;
    mark();
    BK_TRY(gemm(ctx, false, false, m, b, m, 1.0, A22, lda, VT.p, m, 0.0, PA + blk + r0, n));  // Z1 = A22 V1 T1
    mark();
    flops += 2.0 * m * (double)m * b;
    BK_TRY(finish_w(r0, PA, PA + blk, PB, T1));
    const int r1 = r0 + b, m2 = n - r1;
    if (!pair_panels || full_update || m2 < 2) {
      // single panel: A22 -= [V W][W V]' (lower tiles computed, mirrored by the epilogue)
      mark();
      BK_TRY(gemm(ctx, false, true, m, m, 2 * b, -1.0, PA + r0, n, PB + r0, n, 1.0, A22, lda, full_update ? 0 : 2));
      mark();
      flops += (full_update ? 2.0 : 1.0) * m * (double)m * 2 * b;
      c0 += b;
      k += 1;
      continue;
    }
    // ---- the columns of the second panel get the first panel's update now
    BK_TRY(gemm(ctx, false, true, m, b, 2 * b, -1.0, PA + r0, n, PB + r0, n, 1.0, A22, lda));
    // ---- second panel: V2 -> PA[2], PB[3]; W2 -> PA[3], PB[2]
    double* T2 = Tstore + (size_t)(k + 1) * b * b;
    double* A22b = A + r1 + (long long)r1 * lda;
    BK_TRY(factor_panel(r0, PA + 2 * blk, PB + 3 * blk, T2));
    mark();
    BK_TRY(gemm(ctx, false, false, m2, b, m2, 1.0, A22b, lda, VT.p, m2, 0.0, PA + 3 * blk + r1, n));  // A22 V2 T2
    mark();
    flops += 2.0 * m2 * (double)m2 * b;
    BK_TRY(gemm(ctx, true, false, 2 * b, b, m2, 1.0, PB + r1, n, VT.p, m2, 0.0, Gs.p, 2 * b));          // [W1 V1]' V2 T2
    BK_TRY(gemm(ctx, false, false, m2, b, 2 * b, -1.0, PA + r1, n, Gs.p, 2 * b, 1.0, PA + 3 * blk + r1, n));
    BK_TRY(finish_w(r1, PA + 2 * blk, PA + 3 * blk, PB + 2 * blk, T2));
    const int r2 = r1 + b, m3 = n - r2;
    if (pair_lookahead && m3 >= 2) {
      // ---- the next pair's first panel: its columns (rows r1.., incl. its diagonal block) get the rank-4b update
      // first, then its factorisation runs on the side stream under the update of the trailing (m2-b) block
      BK_TRY(gemm(ctx, false, true, m2, b, 4 * b, -1.0, PA + r1, n, PB + r1, n, 1.0, A22b, lda));
      BK_CUDA(cudaEventRecord(e_col, main_st));
      {
        BK_CUDA(cudaStreamWaitEvent(side_st, e_col, 0));
        SideStreamScope sc(ctx);
        double* PAn = PAbuf.p + (size_t)((pair + 1) & 1) * 4 * blk;
        double* PBn = PBbuf.p + (size_t)((pair + 1) & 1) * 4 * blk;
        BK_TRY(factor_panel(r1, PAn, PBn + blk, Tstore + (size_t)(k + 2) * b * b));
        BK_CUDA(cudaEventRecord(e_fact, side_st));
      }
      ahead = true;
      mark();
      BK_TRY(gemm(ctx, false, true, m3, m3, 4 * b, -1.0, PA + r2, n, PB + r2, n, 1.0, A + r2 + (long long)r2 * lda, lda, 2));
      mark();
      flops += 1.0 * m3 * (double)m3 * 4 * b;
    } else {
      // ---- one rank-4b update for both panels
      mark();
      BK_TRY(gemm(ctx, false, true, m2, m2, 4 * b, -1.0, PA + r1, n, PB + r1, n, 1.0, A22b, lda, 2));
      mark();
      flops += 1.0 * m2 * (double)m2 * 4 * b;
    }
    c0 += 2 * b;
    k += 2;
  }
  if (ahead) BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
  extract_band_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n * ldab, 256), 16LL * ctx->sm_count), 256, 0,
                        ctx->stream>>>(A, lda, n, b, AB, ldab);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  if (prof.p) {
    long long h[8];
    BK_CUDA(cudaMemcpyAsync(h, prof.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    BK_CUDA(cudaStreamSynchronize(ctx->stream));
    const double nc = (double)std::max(1LL, h[4]);
    fprintf(stderr, "[panel qr prof, CTA 0] columns %lld: cycles per column: partial dots %.0f, barrier %.0f, reduction %.0f, update %.0f\n",
            h[4], h[0] / nc, h[1] / nc, h[2] / nc, h[3] / nc);
  }
  if (stats) {
    // time during which at least one of the large GEMMs ran (in the pipelined phase products and updates overlap on two
    // streams: their intervals are merged, not added)
    std::vector<std::pair<double, double>> iv;
    for (size_t i = ev0; i + 1 < n_ev; i += 2) {
      float s0 = 0.f, s1 = 0.f;
      cudaEventElapsedTime(&s0, ctx->event_pool[ev0], ctx->event_pool[i]);
      cudaEventElapsedTime(&s1, ctx->event_pool[ev0], ctx->event_pool[i + 1]);
      iv.emplace_back((double)s0, (double)s1);
    }
    std::sort(iv.begin(), iv.end());
    double sec = 0.0, cur_s = 0.0, cur_e = -1.0;
    for (auto& x : iv) {
      if (cur_e < 0.0) {
        cur_s = x.first;
        cur_e = x.second;
      } else if (x.first <= cur_e) {
        cur_e = std::max(cur_e, x.second);
      } else {
        sec += cur_e - cur_s;
        cur_s = x.first;
        cur_e = x.second;
      }
    }
    if (cur_e >= 0.0) sec += cur_e - cur_s;
    stats->gemm_launches = (double)((n_ev - ev0) / 2);
    stats->gemm_seconds = sec * 1e-3;
    stats->gemm_flops = flops;
  }
  return BK_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Distributed dense -> band over the ranks of a peer communicator (SURVEY 8 f4 / e-S2).
//
// The trailing matrix is distributed by COLUMN BLOCKS of width b, block-cyclically: global block J lives on rank
// J mod G as local block J / G (full columns, so the symmetric matrix is stored twice across the machine; every
// rank then updates only its own columns and no reduction is needed).  Per panel J (owner o = J mod G):
//
//   o      : Householder QR of the panel (cooperative kernel above), T from V'V
//   o -> * : the factored panel (R, V, rows c0..n) and T are STORED INTO EVERY OTHER RANK'S HBM over NVLink
//            (peer_push2d) and a flag is raised                                             [b (m+b) doubles]
//   all    : V T;   Z_g = A[:, own active columns]' (V T)   - the rows of Z = A22 V T that this rank owns;
//            scattered into every rank's Z buffer by the same kind of peer stores           [all-gather, m b doubles]
//   all    : W = Z - 1/2 V (T' (V'Z))  (replicated, tiny),   A[:, own columns] -= V W_g' + W V_g'
//
// Everything is ordered by system-scope flags inside kernels on the library stream: no host synchronisation, no
// NCCL call.  Every rank also deposits every factored panel into its full-size matrix `Afact`, so that at the end
// each rank holds exactly what the single-GPU sy2sb leaves behind (band + reflectors + T factors) and the
// band -> tridiagonal stage and the back-transformations run unchanged.
// Work per rank: 6 m^2 b / G flops per panel against 4 m^2 b on one GPU (the symmetry saving of the update is
// traded for independence).
// ---------------------------------------------------------------------------------------------------------
__global__ void dist_unpack_panel_kernel(const double* __restrict__ src, long long lds, int n, int c0, int wcols,
                                         double* __restrict__ Afact, long long lda, double* __restrict__ V,
                                         double* __restrict__ V2, long long ldv, bool write_v) {
  // src: rows c0..n-1 (row index relative to c0), SB columns.  Afact[c0.., c0..c0+SB) = src; explicit V for rows >= r0
  const int rows = n - c0;
  const int m = rows - SB, nr = min(SB, m - 1);
  const long long total = (long long)rows * wcols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows), l = (int)(idx / rows);
    const double x = src[r + (long long)l * lds];
    Afact[(long long)(c0 + r) + (long long)(c0 + l) * lda] = x;
    if (write_v && r >= SB) {
      const int pr = r - SB;  // panel row
      double v = 0.0;
      if (l < nr) v = (pr > l) ? x : (pr == l ? 1.0 : 0.0);
      V[(long long)(c0 + r) + (long long)l * ldv] = v;
      V2[(long long)(c0 + r) + (long long)l * ldv] = v;
    }
  }
}

// Zloc (nact x SB, local column order) -> rows of every rank's Z buffer (global row = global column index), flag
__global__ void __launch_bounds__(256) dist_zpush_kernel(PeerDev pd, const double* __restrict__ Zloc, int nact, int lc0,
                                                         size_t z_off, long long ldz, unsigned dst_mask, unsigned seq) {
  const long long total = (long long)nact * SB;
  for (int r = 0; r < pd.world; ++r) {
    if (!(dst_mask & (1u << r))) continue;
    double* Z = reinterpret_cast<double*>(pd.heap[r] + z_off);
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x) {
      const int i = (int)(idx % nact), j = (int)(idx / nact);
      const int lc = lc0 + i;
      const long long grow = ((long long)(lc / SB) * pd.world + pd.rank) * SB + (lc % SB);
      Z[grow + (long long)j * ldz] = Zloc[i + (long long)j * nact];
    }
  }
  peer_signal_last_cta(pd, peer_counter(pd, CH_ZGATHER), gridDim.x, dst_mask, CH_ZGATHER, seq);
}

// dst (nact x cols) = rows of src (ld lds) at the global indices of this rank's local columns lc0 .. lc0+nact-1
__global__ void dist_gather_rows_kernel(const double* __restrict__ src, long long lds, int nact, int lc0, int cols,
                                        int rank, int world, double* __restrict__ dst) {
  const long long total = (long long)nact * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nact), j = (int)(idx / nact);
    const int lc = lc0 + i;
    const long long grow = ((long long)(lc / SB) * world + rank) * SB + (lc % SB);
    dst[i + (long long)j * nact] = src[grow + (long long)j * lds];
  }
}

// Xg (ncl x p) = rows of X at this rank's global columns
__global__ void dist_gather_x_kernel(const double* __restrict__ X, long long ldx, int p, int ncl, int rank, int world,
                                     double* __restrict__ Xg) {
  const long long total = (long long)ncl * p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % ncl), j = (int)(idx / ncl);
    const long long grow = ((long long)(i / SB) * world + rank) * SB + (i % SB);
    Xg[i + (long long)j * ncl] = X[grow + (long long)j * ldx];
  }
}

// number of valid local columns of `rank` for an n x n matrix
static int dist_local_cols(int n, int rank, int world) {
  const int nb = (int)ceil_div(n, SB);
  int cols = 0;
  for (int J = rank; J < nb; J += world) cols += std::min(SB, n - J * SB);
  return cols;
}

size_t sy2sb_dist_heap_bytes(int n) {
  // two panel receive buffers (n x b + T), two Z buffers (n x b)
  return sizeof(double) * (2 * ((size_t)n * SB + SB * SB + 256) + 2 * (size_t)n * SB) + 4096;
}

int sy2sb_dist(bk_ctx* ctx, bk_peer* peer, const double* X, long long ldx, int p, double sigma, int n, double* Afact,
               double* Tstore, double* AB, int ldab, DevBuf<double>& aloc_cache, BandStats* stats) {
  const int b = SB, G = peer->world, g = peer->rank;
  BK_REQUIRE(ldab >= 2 * b, "sy2sb: band storage needs 2b rows");
  cudaStream_t st = ctx->stream;
  const int ncl = dist_local_cols(n, g, G);
  const long long lda = n;
  const unsigned all = (1u << G) - 1u, others = all & ~(1u << g);

  // ---- symmetric buffers -------------------------------------------------------------------------------
  size_t vrecv_off[2], z_off[2];
  const size_t panel_elems = (size_t)n * b;
  for (int q = 0; q < 2; ++q) BK_TRY(peer_alloc(peer, sizeof(double) * (panel_elems + b * b + 256), &vrecv_off[q]));
  for (int q = 0; q < 2; ++q) BK_TRY(peer_alloc(peer, sizeof(double) * panel_elems, &z_off[q]));

  // ---- local buffers -----------------------------------------------------------------------------------
  DevBuf<double> Aloc, Xg, taus, part, prow, S, VT, S2, S3, PAbuf, PBbuf, Zloc, PBc;
  BK_TRY(Aloc.borrow(aloc_cache, (size_t)n * std::max(1, ncl)));
  BK_TRY(Xg.alloc((size_t)std::max(1, ncl) * p));
  BK_TRY(taus.alloc(b));
  BK_TRY(part.alloc((size_t)2 * ctx->sm_count * qr_ctas_per_sm() * b));
  BK_TRY(prow.alloc(2 * b));
  BK_TRY(S.alloc(b * b));
  BK_TRY(VT.alloc((size_t)n * b));
  BK_TRY(S2.alloc(b * b));
  BK_TRY(S3.alloc(b * b));
  BK_TRY(PAbuf.borrow(ctx->panel_cache[0], (size_t)4 * b * n));  // [V | W] x 2 panel parities
  BK_TRY(PBbuf.borrow(ctx->panel_cache[1], (size_t)4 * b * n));  // [W | V] x 2
  BK_TRY(Zloc.alloc((size_t)std::max(1, ncl) * b));
  BK_TRY(PBc.alloc((size_t)std::max(1, ncl) * 2 * b));
  BK_TRY(ctx->barrier.ensure(4));
  BK_CUDA(cudaMemsetAsync(PAbuf.p, 0, sizeof(double) * (size_t)4 * b * n, st));
  BK_CUDA(cudaMemsetAsync(PBbuf.p, 0, sizeof(double) * (size_t)4 * b * n, st));
  const size_t blk = (size_t)b * n;

  static const int rows_target = getenv("BK_QR_ROWS") ? atoi(getenv("BK_QR_ROWS")) : 128;
  const int max_rows_per = std::max(rows_target, (int)ceil_div(std::max(1, n - b), (int64_t)ctx->sm_count * qr_ctas_per_sm()));
  const size_t max_smem = (size_t)max_rows_per * (b + 1) * sizeof(double);
  BK_REQUIRE(max_smem <= 200 * 1024, "sy2sb: n too large for the shared-memory panel slabs");
  BK_CUDA(cudaFuncSetAttribute(panel_qr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)max_smem));
  BK_CUDA(cudaFuncSetAttribute(sb_larft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)(sizeof(double) * 2 * b * b)));

  // ---- own columns of K straight from X (no gather of the kernel matrix) ---------------------------------
  if (ncl > 0) {
    dist_gather_x_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)ncl * p, 256), 8LL * ctx->sm_count), 256, 0,
                           st>>>(X, ldx, p, ncl, g, G, Xg.p);
    BK_LAUNCHED(ctx);
    BK_TRY(gauss_kernel_rect(ctx, X, ldx, n, Xg.p, ncl, ncl, p, sigma, Aloc.p, lda));
  }
  // everyone's receive buffers are free (previous fit finished everywhere) before the first push
  BK_TRY(peer_barrier(peer, st));

  double flops = 0.0;
  int launches = 0;
  const int nbk = (int)ceil_div(n, b);
  // Sequence numbers of the panel channel are a fixed function of the panel index (two per panel: T, panel), so that
  // the owner of panel J+1 can push it from the side stream while panel J's update is still running everywhere.
  const unsigned seq_base = peer->seq[CH_PANEL];
  peer->seq[CH_PANEL] = seq_base + 2u * (unsigned)nbk;
  auto seqT_of = [&](int J) { return seq_base + 2u * (unsigned)J + 1u; };
  auto seqP_of = [&](int J) { return seq_base + 2u * (unsigned)J + 2u; };
  static const bool lookahead = getenv("BK_DIST_NO_LOOKAHEAD") == nullptr;
  cudaStream_t main_st = ctx->stream, side_st = ctx->side_stream;
  cudaEvent_t e_col = pool_event(ctx, 0), e_fact = pool_event(ctx, 1);

  // Owner side of panel J on the CURRENT ctx->stream: QR in the local columns, T, pushes, own deposit.
  // PAq / PBq: the [V W] / [W V] buffers of this panel's parity.
  auto owner_factor = [&](int J, double* PAq, double* PBq) -> int {
    const int c0 = J * b, r0 = c0 + b, m = n - r0, lb = J / G, q = J & 1;
    cudaStream_t cs = ctx->stream;
    double* Tk = Tstore + (size_t)J * b * b;
    PanelArgs pa;
    pa.A = Aloc.p + ((long long)lb * b - c0) * lda;  // so that A[(r0+r) + (c0+l) lda] is the local column
    pa.lda = lda;
    pa.n = n;
    pa.c0 = c0;
    pa.V = PAq;
    pa.V2 = PBq + blk;
    pa.ldv = n;
    pa.taus = taus.p;
    pa.part = part.p;
    pa.prow = prow.p;
    pa.barrier = ctx->barrier.p;
    pa.prof = nullptr;
    const int Gp = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)ctx->sm_count * qr_ctas_per_sm(), ceil_div(m, rows_target)));
    pa.rows_per = (int)ceil_div(m, Gp);
    BK_CUDA(cudaMemsetAsync(ctx->barrier.p, 0, sizeof(unsigned) * 4, cs));
    void* kargs[] = {&pa};
    const size_t smem = (size_t)pa.rows_per * (b + 1) * sizeof(double);
    BK_CUDA(cudaLaunchCooperativeKernel((void*)panel_qr_kernel, dim3(Gp), dim3(QR_NT), kargs, smem, cs));
    BK_LAUNCHED(ctx);
    BK_TRY(gemm(ctx, true, false, b, b, m, 1.0, PAq + r0, n, PAq + r0, n, 0.0, S.p, b));
    launch_larft(cs, S.p, taus.p, b, Tk);
    BK_LAUNCHED(ctx);
    // T, then the factored panel (rows c0..n: diagonal block, R, V), into every other rank's HBM
    const double* pan = Aloc.p + c0 + (long long)lb * b * lda;
    if (others) {
      BK_TRY(peer_push2d(peer, Tk, b, b, b, vrecv_off[q] + sizeof(double) * panel_elems, b, others, CH_PANEL, seqT_of(J), cs));
      BK_TRY(peer_push2d(peer, pan, lda, n - c0, b, vrecv_off[q] + sizeof(double) * c0, n, others, CH_PANEL, seqP_of(J), cs));
    }
    // own deposit (V is already explicit in PA / PB from the QR kernel)
    dist_unpack_panel_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)(n - c0) * b, 256), 4LL * ctx->sm_count),
                               256, 0, cs>>>(pan, lda, n, c0, b, Afact, n, PAq, PBq + blk, n, false);
    BK_LAUNCHED(ctx);
    return BK_OK;
  };

  const bool prof = getenv("BK_DIST_PROF") != nullptr;
  size_t n_pev = 2;
  auto pmark = [&]() {
    if (prof) cudaEventRecord(pool_event(ctx, n_pev++), main_st);
  };
  int J = 0;
  bool factored_ahead = false;  // this rank owns the current panel and factored it on the side stream already
  for (; J < nbk; ++J) {
    const int c0 = J * b, r0 = c0 + b, m = n - r0;
    if (m < 2) break;
    const int owner = J % G, q = J & 1;
    pmark();  // 0: panel start
    double* PA = PAbuf.p + (size_t)q * 2 * blk;  // [V | W]
    double* PB = PBbuf.p + (size_t)q * 2 * blk;  // [W | V]
    double* Tk = Tstore + (size_t)J * b * b;
    double* vrecv = peer_ptr(peer, vrecv_off[q]);
    if (g == owner) {
      if (factored_ahead)
        BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
      else
        BK_TRY(owner_factor(J, PA, PB));
      factored_ahead = false;
    } else {
      BK_TRY(peer_wait(peer, CH_PANEL, 1u << owner, seqP_of(J), st));
      dist_unpack_panel_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)(n - c0) * b, 256), 4LL * ctx->sm_count),
                                 256, 0, st>>>(vrecv + c0, n, n, c0, b, Afact, n, PA, PB + blk, n, true);
      BK_LAUNCHED(ctx);
      BK_CUDA(cudaMemcpyAsync(Tk, vrecv + panel_elems, sizeof(double) * b * b, cudaMemcpyDeviceToDevice, st));
    }
    pmark();  // 1: panel available (owner: factored / joined; others: received + unpacked)
    BK_TRY(gemm(ctx, false, false, m, b, b, 1.0, PA + r0, n, Tk, b, 0.0, VT.p, m));  // V T
    pmark();  // 2: V T
    // ---- own rows of Z = A22 (V T): the local active columns, transposed
    const int lb0 = (J >= g) ? (J - g) / G + 1 : 0;  // local blocks with global index <= J are done
    const int lc0 = lb0 * b;
    const int nact = std::max(0, ncl - lc0);
    const unsigned seqZ = peer_next_seq(peer, CH_ZGATHER);
    if (nact > 0) {
      BK_TRY(gemm(ctx, true, false, nact, b, m, 1.0, Aloc.p + r0 + (long long)lc0 * lda, lda, VT.p, m, 0.0, Zloc.p, nact));
      flops += 2.0 * nact * (double)m * b;
      ++launches;
    }
    pmark();  // 3: own rows of Z
    {
      const long long tot = (long long)std::max(1, nact) * b;
      dist_zpush_kernel<<<(unsigned)std::min<long long>(ceil_div(tot, 512), 2LL * ctx->sm_count), 256, 0, st>>>(
          peer->dev, Zloc.p, nact, lc0, z_off[q], n, all, seqZ);
      BK_LAUNCHED(ctx);
    }
    BK_TRY(peer_wait(peer, CH_ZGATHER, all, seqZ, st));
    pmark();  // 4: Z gathered
    // ---- W = Z - 1/2 V (T' (V'Z))  (replicated)
    const double* Zfull = peer_ptr(peer, z_off[q]);
    BK_TRY(copy_matrix(ctx, Zfull + r0, n, m, b, 1.0, PA + blk + r0, n));
    BK_TRY(finish_w_fused(ctx, n, r0, PA, PA + blk, PB, Tk, S2.p, S3.p));
    pmark();  // 5: W
    // ---- A[:, own active columns] -= [V W] [W_g V_g]'
    if (nact > 0) {
      const long long tot = (long long)nact * 2 * b;
      dist_gather_rows_kernel<<<(unsigned)std::min<long long>(ceil_div(tot, 256), 4LL * ctx->sm_count), 256, 0, st>>>(
          PB, n, nact, lc0, 2 * b, g, G, PBc.p);
      BK_LAUNCHED(ctx);
      double* C = Aloc.p + r0 + (long long)lc0 * lda;
      const bool next_is_mine = lookahead && ((J + 1) % G == g) && (n - (r0 + b) >= 2) && nact >= b;
      if (next_is_mine) {
        // the next panel is this rank's first active block: update it first, factor it on the side stream under
        // the rest of the update (QR, T and the pushes of panel J+1 leave the critical path of every rank)
        BK_TRY(gemm(ctx, false, true, m, b, 2 * b, -1.0, PA + r0, n, PBc.p, nact, 1.0, C, lda));
        BK_CUDA(cudaEventRecord(e_col, main_st));
        {
          BK_CUDA(cudaStreamWaitEvent(side_st, e_col, 0));
          SideStreamScope sc(ctx);
          BK_TRY(owner_factor(J + 1, PAbuf.p + (size_t)(q ^ 1) * 2 * blk, PBbuf.p + (size_t)(q ^ 1) * 2 * blk));
          BK_CUDA(cudaEventRecord(e_fact, side_st));
        }
        factored_ahead = true;
        if (nact > b)
          BK_TRY(gemm(ctx, false, true, m, nact - b, 2 * b, -1.0, PA + r0, n, PBc.p + b, nact, 1.0, C + (long long)b * lda, lda));
      } else {
        BK_TRY(gemm(ctx, false, true, m, nact, 2 * b, -1.0, PA + r0, n, PBc.p, nact, 1.0, C, lda));
      }
      flops += 2.0 * m * (double)nact * 2 * b;
    }
    pmark();  // 6: update queued behind
  }
  if (factored_ahead) BK_CUDA(cudaStreamWaitEvent(main_st, e_fact, 0));
  const int n_panels = J;
  double* PA = PAbuf.p;
  double* PB = PBbuf.p;
  // ---- the remaining (unfactored) diagonal blocks reach everyone the same way
  for (; J < nbk; ++J) {
    const int c0 = J * b, owner = J % G, lb = J / G, q = J & 1;
    const int w = std::min(b, n - c0);
    const unsigned seqP = seqP_of(J);
    double* vrecv = peer_ptr(peer, vrecv_off[q]);
    const long long tot = (long long)(n - c0) * w;
    const unsigned nblk = (unsigned)std::min<long long>(ceil_div(tot, 256), 4LL * ctx->sm_count);
    if (g == owner) {
      const double* pan = Aloc.p + c0 + (long long)lb * b * lda;
      if (others) BK_TRY(peer_push2d(peer, pan, lda, n - c0, w, vrecv_off[q] + sizeof(double) * c0, n, others, CH_PANEL, seqP, st));
      dist_unpack_panel_kernel<<<nblk, 256, 0, st>>>(pan, lda, n, c0, w, Afact, n, PA, PB + blk, n, false);
      BK_LAUNCHED(ctx);
    } else {
      BK_TRY(peer_wait(peer, CH_PANEL, 1u << owner, seqP, st));
      dist_unpack_panel_kernel<<<nblk, 256, 0, st>>>(vrecv + c0, n, n, c0, w, Afact, n, PA, PB + blk, n, false);
      BK_LAUNCHED(ctx);
    }
  }
  extract_band_kernel<<<(unsigned)std::min<long long>(ceil_div((long long)n * ldab, 256), 16LL * ctx->sm_count), 256, 0,
                        st>>>(Afact, n, n, b, AB, ldab);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  // nobody re-uses (resets) the symmetric buffers before every rank is through
  BK_TRY(peer_barrier(peer, st));
  BK_TRY(peer_check(peer, st));
  if (prof && n_panels > 0) {
    // mean seconds per segment on this rank, panels it owns vs panels it receives
    const char* seg[6] = {"panel ready", "V T", "Z rows (GEMM)", "Z gather", "W", "update"};
    double own[6] = {0}, oth[6] = {0};
    int n_own = 0, n_oth = 0;
    for (int Jp = 0; Jp < n_panels; ++Jp) {
      const bool mine = (Jp % G) == g;
      (mine ? n_own : n_oth)++;
      for (int i = 0; i < 6; ++i) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ctx->event_pool[2 + 7 * Jp + i], ctx->event_pool[2 + 7 * Jp + i + 1]);
        (mine ? own : oth)[i] += ms * 1e-3;
      }
    }
    fprintf(stderr, "[sy2sb_dist prof rank %d/%d, n=%d] total seconds by segment (owned %d panels | received %d panels):\n", g, G, n,
            n_own, n_oth);
    for (int i = 0; i < 6; ++i) fprintf(stderr, "    %-16s %.4f | %.4f\n", seg[i], own[i], oth[i]);
  }
  if (stats) {
    // per-rank flops of the Z products and updates; the launches are not bracketed by events in this variant
    stats->gemm_launches = launches;
    stats->gemm_flops = flops;
    stats->gemm_seconds = 0.0;
  }
  return BK_OK;
}

int sy2sb_bandwidth() { return SB; }

}  // namespace bk
