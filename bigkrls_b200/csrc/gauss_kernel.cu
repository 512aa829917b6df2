// Gaussian kernel build, square-symmetric and rectangular.
//
// Replaces reference src/gauss_kernel.cpp:18-23 (K[i,j] = exp(-sum_d (x_id-x_jd)^2 / sigma),
// j >= i then mirrored) and src/temp_kernel.cpp:19-23 (rectangular, predict()).
//
// Layout: X is n x p column-major (ld = ldx), K column-major.  One CTA produces a 64 x 64
// tile: the two 64 x p row panels of X are staged in shared memory (coalesced along the row
// index), every thread accumulates a 4 x 4 block of squared distances by DIRECT DIFFERENCES
// (bit-compatible with the reference's arithmetic up to the final exp; the Gram form
// |xi|^2+|xj|^2-2xi.xj was rejected: the kernel is bound by the 8N^2-byte store and the FP64
// exp, not by the P-long contraction, and direct differences keep K[i,i] == 1 exactly).
// exp is fused; the tile goes through a padded shared-memory buffer so that BOTH the (i,j)
// block and its mirror (j,i) are written with coalesced 512-byte row segments.  Only tiles
// with bj >= bi are launched (triangular block index), so each exp is evaluated once.
//
// Roofline: HBM store bound, 8 N^2 bytes (+ 8 N p read).
#include "common.cuh"
#include "kernels.cuh"

namespace bk {

static constexpr int GT = 64;   // tile edge
static constexpr int GPC = 16;  // dims staged per chunk

template <bool SYM>
__global__ void __launch_bounds__(256)
    gauss_tile_kernel(const double* __restrict__ A, long long lda, int m,
                      const double* __restrict__ B, long long ldb, int n, int p, double sigma,
                      double* __restrict__ out, long long ldo, int tiles_m) {
  // the staged X panels (2 x 16 x 64) and the output tile (64 x 65) share one buffer
  __shared__ __align__(16) double sbuf[GT * (GT + 1)];
  double(*xa)[GT] = reinterpret_cast<double(*)[GT]>(sbuf);
  double(*xb)[GT] = reinterpret_cast<double(*)[GT]>(sbuf + GPC * GT);
  double(*tile)[GT + 1] = reinterpret_cast<double(*)[GT + 1]>(sbuf);

  int bi, bj;
  if (SYM) {
    // linear index over the upper triangle of tiles, column by column: L = bj(bj+1)/2 + bi
    const long long L = blockIdx.x;
    long long c = (long long)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
    while ((c + 1) * (c + 2) / 2 <= L) ++c;
    while (c * (c + 1) / 2 > L) --c;
    bj = (int)c;
    bi = (int)(L - c * (c + 1) / 2);
  } else {
    bi = blockIdx.x % tiles_m;
    bj = blockIdx.x / tiles_m;
  }
  const int i0 = bi * GT, j0 = bj * GT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;

  for (int d0 = 0; d0 < p; d0 += GPC) {
    const int dc = min(GPC, p - d0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < GT * GPC; idx += 256) {
      const int r = idx & (GT - 1), d = idx >> 6;
      double va = 0.0, vb = 0.0;
      if (d < dc) {
        if (i0 + r < m) va = A[(long long)(i0 + r) + (long long)(d0 + d) * lda];
        if (j0 + r < n) vb = B[(long long)(j0 + r) + (long long)(d0 + d) * ldb];
      }
      xa[d][r] = va;
      xb[d][r] = vb;
    }
    __syncthreads();
#pragma unroll 4
    for (int d = 0; d < dc; ++d) {
      double ra[4], rb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ra[a] = xa[d][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) rb[b] = xb[d][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const double df = ra[a] - rb[b];
          acc[a][b] = fma(df, df, acc[a][b]);
        }
    }
  }
  __syncthreads();  // everyone is done with xa/xb before the buffer becomes the tile
  // exp(-d2 / sigma) as exp(d2 * (-1/sigma)): the kernel is ISSUE bound (ncu: 77 % of the issue slots, FP64 pipe
  // 53 %, DRAM 40 %), and an FP64 division is ~20 instructions per element.  The scaled argument differs from the
  // quotient by at most one rounding: |dK| <= K |x| 2^-53 <= 4e-17.
  const double ninv = -1.0 / sigma;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) tile[tx + 16 * a][ty + 16 * b] = exp(acc[a][b] * ninv);
  __syncthreads();

  // Stores: 16 bytes per thread (st.global.v2.f64) when the output allows it - half the store instructions of the
  // scalar form for a kernel that is bound by its 8 N^2-byte store.  Full tiles only; edges take the scalar path.
  const bool full = (i0 + GT <= m) && (j0 + GT <= n);
  const bool vec = full && ((ldo & 1) == 0) && ((((uintptr_t)out) & 15u) == 0);
  if (vec) {
    // (i, j) block: a thread owns two consecutive rows of one column
    for (int idx = threadIdx.x; idx < (GT / 2) * GT; idx += 256) {
      const int r = (idx & (GT / 2 - 1)) * 2, c = idx >> 5;
      *reinterpret_cast<double2*>(out + (long long)(i0 + r) + (long long)(j0 + c) * ldo) =
          make_double2(tile[r][c], tile[r + 1][c]);
    }
    if (SYM && bi != bj) {
      // mirror block (j, i): two consecutive columns of the tile are two consecutive rows of the mirror
      for (int idx = threadIdx.x; idx < (GT / 2) * GT; idx += 256) {
        const int c = (idx & (GT / 2 - 1)) * 2, r = idx >> 5;
        *reinterpret_cast<double2*>(out + (long long)(j0 + c) + (long long)(i0 + r) * ldo) =
            make_double2(tile[r][c], tile[r][c + 1]);
      }
    }
    return;
  }
  // (i, j) block: consecutive threads -> consecutive rows i (contiguous in column-major out)
  for (int idx = threadIdx.x; idx < GT * GT; idx += 256) {
    const int r = idx & (GT - 1), c = idx >> 6;
    if (i0 + r < m && j0 + c < n) out[(long long)(i0 + r) + (long long)(j0 + c) * ldo] = tile[r][c];
  }
  if (SYM && bi != bj) {
    // mirror block (j, i): consecutive threads -> consecutive j
    for (int idx = threadIdx.x; idx < GT * GT; idx += 256) {
      const int c = idx & (GT - 1), r = idx >> 6;
      if (i0 + r < m && j0 + c < n)
        out[(long long)(j0 + c) + (long long)(i0 + r) * ldo] = tile[r][c];
    }
  }
}

int gauss_kernel_sym(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double sigma,
                     double* K, long long ldk) {
  if (n <= 0) return BK_OK;
  const long long T = ceil_div(n, GT);
  const long long blocks = T * (T + 1) / 2;
  gauss_tile_kernel<true><<<(unsigned)blocks, 256, 0, ctx->stream>>>(X, ldx, n, X, ldx, n, p, sigma,
                                                                     K, ldk, (int)T);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int gauss_kernel_rect(bk_ctx* ctx, const double* A, long long lda, int m, const double* B,
                      long long ldb, int n, int p, double sigma, double* out, long long ldo) {
  if (m <= 0 || n <= 0) return BK_OK;
  const long long tm = ceil_div(m, GT), tn = ceil_div(n, GT);
  gauss_tile_kernel<false><<<(unsigned)(tm * tn), 256, 0, ctx->stream>>>(A, lda, m, B, ldb, n, p,
                                                                         sigma, out, ldo, (int)tm);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

}  // namespace bk
