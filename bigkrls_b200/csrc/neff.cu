// Effective sample size from absolute pairwise row correlations of X.
// Reference: src/Neffective.cpp:23-64 - rows de-meaned ACROSS columns and scaled to unit norm,
// r = sum_{i>j} |z_i . z_j|,  Neffective = N (1 - 2r/N^2) + 1.
// Same 64x64 tiling as the kernel build (only tile pairs bj >= bi are launched); per-CTA partial
// sums are reduced in a fixed order so the result is deterministic.
#include "common.cuh"
#include "kernels.cuh"

namespace bk {

__global__ void neff_normalize_kernel(const double* __restrict__ X, long long ldx, int n, int p,
                                      double* __restrict__ Z) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s = 0.0;
  for (int d = 0; d < p; ++d) s += X[i + (long long)d * ldx];
  const double mean = s / p;  // :28
  double ss = 0.0;
  for (int d = 0; d < p; ++d) {
    const double v = X[i + (long long)d * ldx] - mean;
    ss += v * v;              // :42-44
  }
  const double nrm = sqrt(ss);
  for (int d = 0; d < p; ++d) Z[i + (long long)d * n] = (X[i + (long long)d * ldx] - mean) / nrm;
}

__global__ void __launch_bounds__(256)
    neff_tile_kernel(const double* __restrict__ Z, int n, int p, double* __restrict__ partial) {
  __shared__ double xa[16][64];
  __shared__ double xb[16][64];
  __shared__ double red[32];
  const long long L = blockIdx.x;
  long long c = (long long)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
  while ((c + 1) * (c + 2) / 2 <= L) ++c;
  while (c * (c + 1) / 2 > L) --c;
  const int bj = (int)c, bi = (int)(L - c * (c + 1) / 2);  // bi <= bj
  const int i0 = bi * 64, j0 = bj * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  double acc[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) acc[a][b] = 0.0;
  for (int d0 = 0; d0 < p; d0 += 16) {
    const int dc = min(16, p - d0);
    __syncthreads();
    for (int idx = threadIdx.x; idx < 64 * 16; idx += 256) {
      const int r = idx & 63, d = idx >> 6;
      double va = 0.0, vb = 0.0;
      if (d < dc) {
        if (i0 + r < n) va = Z[(long long)(i0 + r) + (long long)(d0 + d) * n];
        if (j0 + r < n) vb = Z[(long long)(j0 + r) + (long long)(d0 + d) * n];
      }
      xa[d][r] = va;
      xb[d][r] = vb;
    }
    __syncthreads();
    for (int d = 0; d < dc; ++d) {
      double ra[4], rb[4];
#pragma unroll
      for (int a = 0; a < 4; ++a) ra[a] = xa[d][tx + 16 * a];
#pragma unroll
      for (int b = 0; b < 4; ++b) rb[b] = xb[d][ty + 16 * b];
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fma(ra[a], rb[b], acc[a][b]);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int a = 0; a < 4; ++a)
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int gi = i0 + tx + 16 * a, gj = j0 + ty + 16 * b;
      if (gi < n && gj < n && gi < gj) s += fabs(acc[a][b]);  // each unordered pair once (:53)
    }
  s = block_sum(s, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

__global__ void neff_final_kernel(const double* __restrict__ partial, long long nb, int n,
                                  double* __restrict__ out) {
  __shared__ double red[32];
  // fixed assignment of partials to threads + fixed-shape tree => deterministic
  double s = 0.0;
  for (long long b = threadIdx.x; b < nb; b += blockDim.x) s += partial[b];
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double nn = (double)n;
    const double mean_abs = 2.0 * s / (nn * nn);  // :60
    out[0] = nn * (1.0 - mean_abs) + 1.0;         // :62
  }
}

int neffective_acf(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double* out_host) {
  DevBuf<double> Z, partial, out;
  const long long T = ceil_div(n, 64);
  const long long nb = T * (T + 1) / 2;
  BK_TRY(Z.alloc((size_t)n * p));
  BK_TRY(partial.alloc((size_t)nb));
  BK_TRY(out.alloc(1));
  neff_normalize_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, ctx->stream>>>(X, ldx, n, p, Z.p);
  BK_LAUNCHED(ctx);
  neff_tile_kernel<<<(unsigned)nb, 256, 0, ctx->stream>>>(Z.p, n, p, partial.p);
  BK_LAUNCHED(ctx);
  neff_final_kernel<<<1, 1024, 0, ctx->stream>>>(partial.p, nb, n, out.p);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  BK_CUDA(cudaMemcpyAsync(out_host, out.p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  return BK_OK;
}

}  // namespace bk
