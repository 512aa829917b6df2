// Blocked Householder tridiagonalisation of a dense symmetric matrix (lower triangle),
// A = Q T Q'.  Stage one of the full eigensolver that replaces arma::eig_sym -> LAPACK dsyevd
// (reference src/eigen.cpp:24).  Written from the published algorithm (Dongarra/Sorensen/
// Hammarling blocked reduction: panel of nb reflectors with deferred rank-2nb update).
//
// One PERSISTENT COOPERATIVE kernel per panel (all CTAs co-resident, software grid barrier);
// per column j of the panel:
//   P1  a = A[j:,j] - V W[j,:]' - W V[j,:]'          rows over all threads, + ||a[j+2:]||^2 partials
//   P2  Householder (beta, tau, v) computed redundantly by every CTA from the partials
//   P3  y = A22 v  with A22 the UN-updated trailing matrix: each 64x64 tile of the LOWER
//       triangle is read ONCE and used for both y_i += T v_j and y_j += T' v_i
//       (4 m^2 bytes instead of 8 m^2: this read is the HBM roofline of the whole stage);
//       partial results go to owned slots (no atomics) and are reduced in P4 in a fixed
//       order => bitwise reproducible.  Also u1 = W'v, u2 = V'v.
//   P4  w = tau (y - V u1 - W u2),  partials of w.v ;   (next P1)  w += -tau/2 (w.v) v
// After the panel: A22 -= [V W][W V]'  (lower tiles only) on the DMMA GEMM.
//
// Layout: A column-major n x n; panel workspace P = [V | W | V] (n x 3nb) so that the rank-2nb
// update is ONE GEMM  A22 -= P[:,0:2nb] * P[:,nb:3nb]'.  Reflector j is stored LAPACK-style in
// A[j+2:, j] (v[j+1] = 1 implicit; A[j+1, j] is left holding alpha), e[j] = beta, d[j] = A[j,j].
#include <cooperative_groups.h>
#include <cstdlib>
#include "common.cuh"
#include "dgemm.cuh"
#include "eigen.cuh"
#include "kernels.cuh"

namespace bk {

static constexpr int TS = 64;  // SYMV tile edge

struct SytrdArgs {
  double* A;
  long long lda;
  int n, j0, nb;
  double* P;       // n x 3nb  [V | W | V]
  double* d;
  double* e;
  double* tau;
  double* part;    // per-CTA partial sums (2 * gridDim)
  double* dots;    // 2 * nb + 1: u1 = W'v, u2 = V'v, and w'[j+1] of the current column
  double* ypart;   // [T][n]   direct partials, owned by (strip bc)
  double* ytpart;  // [TSEG][n] transposed partials, owned by (segment sb)
  unsigned* barrier;
};

__device__ __forceinline__ double ldcg(const double* p) { return __ldcg(p); }
// data written by other CTAs earlier in this kernel is always read through L2
#define LDX(p) __ldcg(p)

__global__ void __launch_bounds__(256, 2) sytrd_panel_kernel(SytrdArgs a) {
  __shared__ double red[32];
  __shared__ double sd[2][4][TS];
  __shared__ double st[4][16][2];
  const int n = a.n, nb = a.nb;
  double* __restrict__ A = a.A;
  const long long lda = a.lda;
  double* V = a.P;
  double* W = a.P + (size_t)nb * n;
  double* V2 = a.P + (size_t)2 * nb * n;
  const int G = gridDim.x;
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int gthreads = G * blockDim.x;
  unsigned epoch = 0;
  const int T = (n + TS - 1) / TS;

  double alpha_prev = 0.0;  // -tau/2 (w.v) of the previous column, applied lazily
  int pending = -1;         // panel column whose w still lacks the alpha*v correction
  for (int c = 0; c < nb; ++c) {
    const int j = a.j0 + c;
    if (j >= n) break;
    // ---------------- P1: finish previous w, update column j ----------------------------------
    // previous column's w: W[r, c-1] += alpha_prev * V[r, c-1]   (rows r > j-1, i.e. r >= j)
    // W[j, c-1] after the correction is needed by every thread: w'[j] was parked in
    // dots[2nb] by P4 (V[j, c-1] == 1), so nobody reads the element another thread rewrites.
    double wjc = 0.0;
    if (c > 0) wjc = ldcg(a.dots + 2 * nb) + alpha_prev;
    pending = -1;
    double ss = 0.0;
    for (int r = j + gid; r < n; r += gthreads) {
      if (c > 0) {
        double* wp = W + (size_t)(c - 1) * n;
        wp[r] = LDX(wp + r) + alpha_prev * LDX(V + (size_t)(c - 1) * n + r);
      }
      double v = A[r + (long long)j * lda];
      for (int q = 0; q < c; ++q) {
        const double wjq = (q == c - 1) ? wjc : ldcg(W + (size_t)q * n + j);
        const double vjq = ldcg(V + (size_t)q * n + j);
        v -= LDX(V + (size_t)q * n + r) * wjq + LDX(W + (size_t)q * n + r) * vjq;
      }
      A[r + (long long)j * lda] = v;
      if (r == j) a.d[j] = v;
      if (r > j + 1) ss += v * v;
    }
    if (j == n - 1) break;  // last diagonal element: nothing to reflect
    ss = block_sum(ss, red);
    if (threadIdx.x == 0) a.part[blockIdx.x] = ss;
    grid_barrier(a.barrier, epoch);

    // ---------------- P2: Householder vector ------------------------------------------------
    double xnorm2 = 0.0;  // same fixed-shape reduction in every CTA => identical value everywhere
    for (int b = threadIdx.x; b < G; b += blockDim.x) xnorm2 += ldcg(a.part + b);
    xnorm2 = block_sum(xnorm2, red);
    const double alpha = ldcg(A + (j + 1) + (long long)j * lda);
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
    if (gid == 0) {
      a.e[j] = beta;
      a.tau[j] = tau;
    }
    for (int r = j + 1 + gid; r < n; r += gthreads) {
      double v;
      if (r == j + 1) {
        // A[j+1, j] keeps alpha: every CTA reads it at the top of this phase, so it must not be
        // overwritten here (beta lives in e[j]; the back-transform uses an explicit 1).
        v = 1.0;
      } else {
        v = ldcg(A + r + (long long)j * lda) * scale;
        A[r + (long long)j * lda] = v;
      }
      V[(size_t)c * n + r] = v;
      V2[(size_t)c * n + r] = v;
    }
    grid_barrier(a.barrier, epoch);

    const double* vcol = V + (size_t)c * n;
    // ---------------- P3: u = [W V]'v  and  y = A22 v ------------------------------------------
    for (int q = blockIdx.x; q < 2 * c; q += G) {
      const double* col = (q < c) ? (W + (size_t)q * n) : (V + (size_t)(q - c) * n);
      double s = 0.0;
      for (int r = j + 1 + threadIdx.x; r < n; r += blockDim.x) s += LDX(col + r) * ldcg(vcol + r);
      s = block_sum(s, red);
      if (threadIdx.x == 0) a.dots[(q < c) ? q : (nb + q - c)] = s;
    }
    {
      const int t0 = (j + 1) / TS;            // first tile row/col touching the trailing matrix
      const int Tm = T - t0;                  // tiles per side
      int S = (int)(((long long)Tm * Tm) / (8LL * G));
      S = max(1, min(8, S));
      const int NSEG = (Tm + S - 1) / S;
      const int rr = threadIdx.x & (TS - 1), cg = threadIdx.x >> 6;  // row in tile, column group
      const int lane = threadIdx.x & 31;
      for (int item = blockIdx.x; item < Tm * NSEG; item += G) {
        const int bc = t0 + item % Tm, sb = item / Tm;
        const int br_lo = max(bc, t0 + sb * S), br_hi = min(T, t0 + sb * S + S);
        if (br_lo >= br_hi) continue;  // uniform per CTA
        const int col0 = bc * TS + cg * 16;
        double vc[16], accT[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
          vc[k] = (col0 + k < n) ? ldcg(vcol + col0 + k) : 0.0;
          accT[k] = 0.0;
        }
        int parity = 0;
        for (int br = br_lo; br < br_hi; ++br) {
          const int row = br * TS + rr;
          const bool rok = row < n;
          const double vr = rok ? ldcg(vcol + row) : 0.0;
          const double* ap = A + row + (long long)col0 * lda;
          double av[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) av[k] = (rok && col0 + k < n) ? ap[(long long)k * lda] : 0.0;
          double dsum = 0.0;
          if (br != bc) {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              dsum = fma(av[k], vc[k], dsum);
              accT[k] = fma(av[k], vr, accT[k]);
            }
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const int col = col0 + k;
              if (row > col) {
                dsum = fma(av[k], vc[k], dsum);
                accT[k] = fma(av[k], vr, accT[k]);
              } else if (row == col) {
                dsum = fma(av[k], vc[k], dsum);
              }
            }
          }
          sd[parity][cg][rr] = dsum;
          __syncthreads();
          if (threadIdx.x < TS) {
            const double s = (sd[parity][0][rr] + sd[parity][1][rr]) +
                             (sd[parity][2][rr] + sd[parity][3][rr]);
            if (row < n) a.ypart[(size_t)bc * n + row] = s;
          }
          parity ^= 1;
        }
        // transposed partials: reduce accT over the 64 rows (2 warps per column group)
#pragma unroll
        for (int k = 0; k < 16; ++k) accT[k] = warp_sum(accT[k]);
        __syncthreads();
        if (lane == 0) {
#pragma unroll
          for (int k = 0; k < 16; ++k) st[cg][k][(threadIdx.x >> 5) & 1] = accT[k];
        }
        __syncthreads();
        if (threadIdx.x < 64) {
          const int g4 = threadIdx.x >> 4, k = threadIdx.x & 15;
          const int col = bc * TS + g4 * 16 + k;
          if (col < n) a.ytpart[(size_t)sb * n + col] = st[g4][k][0] + st[g4][k][1];
        }
        __syncthreads();
      }
      grid_barrier(a.barrier, epoch);

      // -------------- P4: w = tau (y - V u1 - W u2), partial w.v ------------------------------
      double wv = 0.0;
      for (int r = j + 1 + gid; r < n; r += gthreads) {
        const int br = r / TS;
        double y = 0.0;
        for (int bcx = t0; bcx <= br; ++bcx) y += ldcg(a.ypart + (size_t)bcx * n + r);
        for (int sbx = (br - t0) / S; sbx < NSEG; ++sbx) y += ldcg(a.ytpart + (size_t)sbx * n + r);
        for (int q = 0; q < c; ++q)
          y -= LDX(V + (size_t)q * n + r) * ldcg(a.dots + q) + LDX(W + (size_t)q * n + r) * ldcg(a.dots + nb + q);
        const double w = tau * y;
        W[(size_t)c * n + r] = w;
        if (r == j + 1) a.dots[2 * nb] = w;
        wv += w * ldcg(vcol + r);
      }
      wv = block_sum(wv, red);
      if (threadIdx.x == 0) a.part[G + blockIdx.x] = wv;
      grid_barrier(a.barrier, epoch);
      double tot = 0.0;
      for (int b = threadIdx.x; b < G; b += blockDim.x) tot += ldcg(a.part + G + b);
      tot = block_sum(tot, red);
      alpha_prev = -0.5 * tau * tot;
      pending = c;
    }
  }
  // finish the last w of the panel
  if (pending >= 0) {
    const int c = pending, j = a.j0 + c;
    for (int r = j + 1 + gid; r < n; r += gthreads)
      W[(size_t)c * n + r] += alpha_prev * V[(size_t)c * n + r];
  }
}

int sytrd_lower(bk_ctx* ctx, double* A, long long lda, int n, double* d, double* e, double* tau,
                int nb, SytrdStats* stats) {
  if (n <= 0) return BK_OK;
  if (n == 1) {
    BK_CUDA(cudaMemcpyAsync(d, A, sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
    return BK_OK;
  }
  int occ = 0;
  BK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sytrd_panel_kernel, 256, 0));
  BK_REQUIRE(occ >= 1, "sytrd: kernel does not fit on an SM");
  occ = std::min(occ, 2);
  if (const char* ev = getenv("BK_SYTRD_OCC")) occ = std::max(1, std::min(occ, atoi(ev)));
  const int G = ctx->sm_count * occ;
  const int T = (int)ceil_div(n, TS);
  const int TSEG = T;  // upper bound on the number of row segments (S >= 1)
  DevBuf<double> P, part, dots, ypart, ytpart;
  BK_TRY(P.alloc((size_t)3 * nb * n));
  BK_TRY(part.alloc((size_t)2 * G));
  BK_TRY(dots.alloc((size_t)2 * nb + 1));
  BK_TRY(ypart.alloc((size_t)T * n));
  BK_TRY(ytpart.alloc((size_t)TSEG * n));
  BK_TRY(ctx->barrier.ensure(4));

  // per-panel CUDA events on the launching stream: the roofline numbers of bench.py are the
  // sum of these kernel durations against the algorithmic bytes 4 m_j^2 per column
  const int npanels = (int)ceil_div(n, nb);
  std::vector<cudaEvent_t> ev0(stats ? npanels : 0), ev1(stats ? npanels : 0);
  for (auto& x : ev0) BK_CUDA(cudaEventCreate(&x));
  for (auto& x : ev1) BK_CUDA(cudaEventCreate(&x));
  int pi = 0;
  for (int j0 = 0; j0 < n; j0 += nb, ++pi) {
    BK_CUDA(cudaMemsetAsync(P.p, 0, sizeof(double) * (size_t)3 * nb * n, ctx->stream));
    BK_CUDA(cudaMemsetAsync(ctx->barrier.p, 0, sizeof(unsigned) * 4, ctx->stream));
    SytrdArgs args;
    args.A = A;
    args.lda = lda;
    args.n = n;
    args.j0 = j0;
    args.nb = nb;
    args.P = P.p;
    args.d = d;
    args.e = e;
    args.tau = tau;
    args.part = part.p;
    args.dots = dots.p;
    args.ypart = ypart.p;
    args.ytpart = ytpart.p;
    args.barrier = ctx->barrier.p;
    void* kargs[] = {&args};
    if (stats) BK_CUDA(cudaEventRecord(ev0[pi], ctx->stream));
    BK_CUDA(cudaLaunchCooperativeKernel((void*)sytrd_panel_kernel, dim3(G), dim3(256), kargs, 0,
                                        ctx->stream));
    if (stats) BK_CUDA(cudaEventRecord(ev1[pi], ctx->stream));
    BK_LAUNCHED(ctx);
    const int jn = j0 + nb;
    if (jn < n) {
      // A22 -= [V W] [W V]'   (lower tiles only)
      const int m = n - jn;
      BK_TRY(gemm(ctx, false, true, m, m, 2 * nb, -1.0, P.p + jn, n, P.p + (size_t)nb * n + jn, n,
                  1.0, A + jn + (long long)jn * lda, lda, true));
    }
  }
  BK_CUDA(cudaStreamSynchronize(ctx->stream));  // workspaces are freed on return
  if (stats) {
    stats->launches = npanels;
    stats->kernel_seconds = 0.0;
    stats->algorithmic_bytes = 0.0;
    for (int i = 0; i < npanels; ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, ev0[i], ev1[i]);
      stats->kernel_seconds += 1e-3 * ms;
      cudaEventDestroy(ev0[i]);
      cudaEventDestroy(ev1[i]);
    }
    for (int j = 0; j < n - 1; ++j) {
      const double m = (double)(n - 1 - j);
      stats->algorithmic_bytes += 4.0 * m * m;  // lower triangle (incl. diagonal ~ m^2/2 doubles)
    }
  }
  return BK_OK;
}

}  // namespace bk
