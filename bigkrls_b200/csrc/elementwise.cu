// Small bandwidth-bound kernels around the contractions: column scaling (multdiag),
// symmetrisation, the right-hand sides and epilogue of the marginal-effects K-pass, spectral
// quadratic forms.  All matrices column-major; consecutive threads walk consecutive rows so
// every global access is coalesced.
#include "common.cuh"
#include "kernels.cuh"

namespace bk {

static inline int grid_for(long long total, int block, int cap) {
  long long b = (total + block - 1) / block;
  if (b < 1) b = 1;
  if (b > cap) b = cap;
  return (int)b;
}

// ---- col_scale (reference src/multdiag.cpp:17-18) -------------------------------------------
__global__ void col_scale_kernel(const double* __restrict__ A, long long lda, int n, int k,
                                 const double* __restrict__ d, const double* __restrict__ scalar,
                                 double* __restrict__ out, long long ldo) {
  const double sc = scalar ? *scalar : 1.0;
  const long long total = (long long)n * k;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % n), c = (int)(idx / n);
    out[r + (long long)c * ldo] = A[r + (long long)c * lda] * (sc * d[c]);
  }
}
int col_scale(bk_ctx* ctx, const double* A, long long lda, int n, int k, const double* d,
              const double* dev_scalar, double* out, long long ldo) {
  if (n <= 0 || k <= 0) return BK_OK;
  col_scale_kernel<<<grid_for((long long)n * k, 256, 16 * ctx->sm_count), 256, 0, ctx->stream>>>(
      A, lda, n, k, d, dev_scalar, out, ldo);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- symmetrize: upper <- lower, 32x32 tiles through shared memory --------------------------
__global__ void symmetrize_kernel(double* __restrict__ C, long long ldc, int n) {
  __shared__ double t[32][33];
  // triangular tile index (bi >= bj)
  const long long L = blockIdx.x;
  long long c = (long long)((sqrt(8.0 * (double)L + 1.0) - 1.0) * 0.5);
  while ((c + 1) * (c + 2) / 2 <= L) ++c;
  while (c * (c + 1) / 2 > L) --c;
  const int bi = (int)c, bj = (int)(L - c * (c + 1) / 2);  // bi >= bj
  const int r0 = bi * 32, c0 = bj * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int cc = ty; cc < 32; cc += 8) {
    const int r = r0 + tx, col = c0 + cc;
    t[cc][tx] = (r < n && col < n) ? C[r + (long long)col * ldc] : 0.0;
  }
  __syncthreads();
  // write transposed block: element (col, r) <- (r, col), rows of the destination = col index
  for (int rr = ty; rr < 32; rr += 8) {
    const int drow = c0 + tx, dcol = r0 + rr;  // destination (drow, dcol) = source (dcol, drow)
    if (drow < n && dcol < n && dcol > drow) C[drow + (long long)dcol * ldc] = t[tx][rr];
  }
}
int symmetrize_from_lower(bk_ctx* ctx, double* C, long long ldc, int n) {
  if (n <= 1) return BK_OK;
  const long long T = ceil_div(n, 32);
  symmetrize_kernel<<<(unsigned)(T * (T + 1) / 2), 256, 0, ctx->stream>>>(C, ldc, n);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- spectral weights -----------------------------------------------------------------------
__global__ void spectral_weights_kernel(const double* __restrict__ ev, int k, double lam, int mode,
                                        double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x) {
    const double e = ev[i];
    double v;
    if (mode == 0)
      v = 1.0 / (e + lam);
    else if (mode == 1) {
      const double s = e + lam;
      v = 1.0 / (s * s);  // (ev+lambda)^-2, R/bigKRLS.R:299
    } else {
      const double r = e / (e + lam);
      v = r * r;
    }
    out[i] = v;
  }
}
int spectral_weights(bk_ctx* ctx, const double* ev, int k, double lam, int mode, double* out) {
  if (k <= 0) return BK_OK;
  spectral_weights_kernel<<<grid_for(k, 256, 4 * ctx->sm_count), 256, 0, ctx->stream>>>(ev, k, lam,
                                                                                        mode, out);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- binary detection (src/bigderiv_v3.cpp:28-31,34-35) -------------------------------------
__global__ void column_binary_kernel(const double* __restrict__ X, long long ldx, int n,
                                     double* __restrict__ info) {
  __shared__ double red[32];
  const double* x = X + (long long)blockIdx.x * ldx;
  double mn = INFINITY, mx = -INFINITY;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    mn = fmin(mn, v);
    mx = fmax(mx, v);
  }
  // block max of mx and of -mn
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  mx = warp_max(mx);
  double nmn = warp_max(-mn);
  __shared__ double red2[32];
  if (lane == 0) {
    red[wid] = mx;
    red2[wid] = nmn;
  }
  __syncthreads();
  if (wid == 0) {
    double a = (lane < nw) ? red[lane] : -INFINITY;
    double b = (lane < nw) ? red2[lane] : -INFINITY;
    a = warp_max(a);
    b = warp_max(b);
    if (lane == 0) {
      red[0] = a;
      red2[0] = b;
    }
  }
  __syncthreads();
  mx = red[0];
  mn = -red2[0];
  __syncthreads();
  double other = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double v = x[i];
    if (v != mn && v != mx) other += 1.0;
  }
  other = block_sum(other, red);
  if (threadIdx.x == 0) {
    info[4 * blockIdx.x + 0] = mn;
    info[4 * blockIdx.x + 1] = mx;
    info[4 * blockIdx.x + 2] = (other == 0.0 && mn != mx) ? 1.0 : 0.0;
  }
}
// info[4j+3] = rank of column j among the binary columns (their extra K-pass columns), info[4p] = their number
__global__ void binary_slots_kernel(double* __restrict__ info, int p) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int s = 0;
  for (int j = 0; j < p; ++j) {
    const bool b = info[4 * j + 2] != 0.0;
    info[4 * j + 3] = b ? (double)s : -1.0;
    s += b ? 1 : 0;
  }
  info[4 * p] = (double)s;
}
int column_binary_info(bk_ctx* ctx, const double* X, long long ldx, int n, int p, double* info, int* nbin_host) {
  *nbin_host = 0;
  if (p <= 0) return BK_OK;
  column_binary_kernel<<<p, 256, 0, ctx->stream>>>(X, ldx, n, info);
  BK_LAUNCHED(ctx);
  binary_slots_kernel<<<1, 32, 0, ctx->stream>>>(info, p);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  double nb = 0.0;
  BK_CUDA(cudaMemcpyAsync(&nb, info + 4 * p, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  *nbin_host = (int)nb;
  return BK_OK;
}

// ---- K-pass right-hand sides ----------------------------------------------------------------
// Binary columns carry BOTH indicator pairs, [x == z1] and [x == z0]: the sums over the z0 group must not be
// formed as (sum over all) - (sum over the z1 group) - for a rare category (a state dummy with one county) the
// z0-group sum seen from a z1 row is ~exp(-dz^2/sigma) ~ 1e-20 of the total and is later multiplied by
// exp(+dz^2/sigma); the subtraction would leave rounding noise times 1e20.
__global__ void build_rhs_kernel(const double* __restrict__ X, long long ldx, int n, int p, int nbin,
                                 const double* __restrict__ c, const double* __restrict__ info,
                                 double* __restrict__ W, long long ldw) {
  const long long total = (long long)n * (p + 1);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % n), j = (int)(idx / n);
    const double ci = c[r];
    if (j == p) {
      W[r] = 1.0;
      W[r + ldw] = ci;
    } else {
      double v = X[r + (long long)j * ldx];
      if (info[4 * j + 2] != 0.0) {
        const int s = (int)info[4 * j + 3];
        const double b0 = (v == info[4 * j + 0]) ? 1.0 : 0.0;
        v = (v == info[4 * j + 1]) ? 1.0 : 0.0;
        W[r + (long long)(2 + 2 * p + s) * ldw] = b0;
        W[r + (long long)(2 + 2 * p + nbin + s) * ldw] = b0 * ci;
      }
      W[r + (long long)(2 + j) * ldw] = v;
      W[r + (long long)(2 + p + j) * ldw] = v * ci;
    }
  }
}
int build_kpass_rhs(bk_ctx* ctx, const double* X, long long ldx, int n, int p, int nbin, const double* c,
                    const double* info, double* W, long long ldw) {
  build_rhs_kernel<<<grid_for((long long)n * (p + 1), 256, 16 * ctx->sm_count), 256, 0,
                     ctx->stream>>>(X, ldx, n, p, nbin, c, info, W, ldw);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- marginal-effects epilogue --------------------------------------------------------------
__global__ void deriv_epilogue_kernel(const double* __restrict__ X, long long ldx, int n, int p, int nbin,
                                      const double* __restrict__ KW, long long ldkw,
                                      const double* __restrict__ info, double sigma,
                                      double* __restrict__ D, long long ldd,
                                      double* __restrict__ R, long long ldr) {
  const long long total = (long long)n * p;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n), j = (int)(idx / n);
    const double K1 = KW[i], Kc = KW[i + ldkw];
    const double A = KW[i + (long long)(2 + j) * ldkw];
    const double Bc = KW[i + (long long)(2 + p + j) * ldkw];
    const double x = X[i + (long long)j * ldx];
    double d, r;
    if (info[4 * j + 2] == 0.0) {
      // continuous, src/bigderiv_v3.cpp:90-106:  L = (x_kj - x_ij) o K ; D = (-2/sigma) L c
      d = (-2.0 / sigma) * (x * Kc - Bc);
      r = x * K1 - A;
    } else {
      // binary, src/bigderiv_v3.cpp:31-87
      const double z0 = info[4 * j + 0], z1 = info[4 * j + 1];
      const int sl = (int)info[4 * j + 3];
      const double sd = 1.0 / (z1 - z0);                  // :36
      const double phi = -1.0 / (sd * sd * sigma);        // :37
      const double dz = z1 - z0;
      const double e1 = exp(-(dz * dz) / sigma);          // c2 when both rows share the value (:69)
      const double e2 = exp((dz * dz) / sigma);           // c2 otherwise
      const double ep = exp(phi), em = exp(-phi);
      const double S1 = A, C1 = Bc;   // sums over the z1 group; the z0 group has its own columns (see build_rhs_kernel)
      const double S0 = KW[i + (long long)(2 + 2 * p + sl) * ldkw];
      const double C0 = KW[i + (long long)(2 + 2 * p + nbin + sl) * ldkw];
      if (x == z0) {
        d = -sd * ((1.0 - e1) * C0 + (1.0 - e2) * C1);
        r = (ep - 1.0) * S0 + (1.0 - em) * S1;
      } else {
        d = sd * ((1.0 - e2) * C0 + (1.0 - e1) * C1);
        r = (em - 1.0) * S0 + (1.0 - ep) * S1;
      }
    }
    D[i + (long long)j * ldd] = d;
    R[i + (long long)j * ldr] = r;
  }
}
int deriv_epilogue(bk_ctx* ctx, const double* X, long long ldx, int n, int p, int nbin, const double* KW,
                   long long ldkw, const double* info, double sigma, double* D, long long ldd,
                   double* R, long long ldr) {
  deriv_epilogue_kernel<<<grid_for((long long)n * p, 256, 16 * ctx->sm_count), 256, 0,
                          ctx->stream>>>(X, ldx, n, p, nbin, KW, ldkw, info, sigma, D, ldd, R, ldr);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

__device__ __forceinline__ double var_factor(const double* info, int j, double sigma, int n) {
  const double nn = (double)n * (double)n;
  if (info[4 * j + 2] == 0.0) return (1.0 / nn) * ((-2.0 / sigma) * (-2.0 / sigma));  // :105
  const double sd = 1.0 / (info[4 * j + 1] - info[4 * j + 0]);
  return 2.0 * sd * sd / nn;                                                           // :85
}

__global__ void deriv_var_spectral_kernel(const double* __restrict__ G, long long ldg, int k,
                                          const double* __restrict__ w2,
                                          const double* __restrict__ sigmasq,
                                          const double* __restrict__ info, double sigma, int n,
                                          double* __restrict__ var) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const double g = G[i + (long long)j * ldg];
    s += w2[i] * g * g;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) var[j] = var_factor(info, j, sigma, n) * (sigmasq ? *sigmasq : 1.0) * s;
}
int deriv_variance_spectral(bk_ctx* ctx, const double* G, long long ldg, int k, int p,
                            const double* w2, const double* dev_sigmasq, const double* info,
                            double sigma, int n, double* var) {
  if (p <= 0) return BK_OK;
  deriv_var_spectral_kernel<<<p, 256, 0, ctx->stream>>>(G, ldg, k, w2, dev_sigmasq, info, sigma, n,
                                                        var);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

__global__ void deriv_var_dense_kernel(const double* __restrict__ R, long long ldr,
                                       const double* __restrict__ VR, long long ldvr, int n,
                                       const double* __restrict__ info, double sigma,
                                       double* __restrict__ var) {
  __shared__ double red[32];
  const int j = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    s += R[i + (long long)j * ldr] * VR[i + (long long)j * ldvr];
  s = block_sum(s, red);
  if (threadIdx.x == 0) var[j] = var_factor(info, j, sigma, n) * s;
}
int deriv_variance_dense(bk_ctx* ctx, const double* R, long long ldr, const double* VR,
                         long long ldvr, int n, int p, const double* info, double sigma,
                         double* var) {
  if (p <= 0) return BK_OK;
  deriv_var_dense_kernel<<<p, 512, 0, ctx->stream>>>(R, ldr, VR, ldvr, n, info, sigma, var);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- sigma^2 = ||y - yhat||^2 / n  (R/bigKRLS.R:294) -----------------------------------------
__global__ void residual_kernel(const double* __restrict__ y, const double* __restrict__ yhat,
                                int n, double* __restrict__ out) {
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double r = y[i] - yhat[i];
    s += r * r;
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = s / (double)n;
}
int residual_sigmasq(bk_ctx* ctx, const double* y, const double* yhat, int n, double* out) {
  residual_kernel<<<1, 1024, 0, ctx->stream>>>(y, yhat, n, out);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- diag(G diag(s) G') ---------------------------------------------------------------------
__global__ void row_quadform_kernel(const double* __restrict__ G, long long ldg, int m, int k,
                                    const double* __restrict__ s, const double* __restrict__ scalar,
                                    double host_scale, double* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= m) return;
  double acc = 0.0;
  for (int i = 0; i < k; ++i) {
    const double g = G[r + (long long)i * ldg];
    acc = fma(s[i] * g, g, acc);
  }
  out[r] = acc * host_scale * (scalar ? *scalar : 1.0);
}
int row_quadform(bk_ctx* ctx, const double* G, long long ldg, int m, int k, const double* s,
                 const double* dev_scalar, double host_scale, double* out) {
  if (m <= 0) return BK_OK;
  row_quadform_kernel<<<(unsigned)ceil_div(m, 128), 128, 0, ctx->stream>>>(G, ldg, m, k, s,
                                                                           dev_scalar, host_scale,
                                                                           out);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

// ---- copies ---------------------------------------------------------------------------------
__global__ void copy_matrix_kernel(const double* __restrict__ src, long long lds, int rows,
                                   int cols, double alpha, double* __restrict__ dst,
                                   long long ldd) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows), c = (int)(idx / rows);
    dst[r + (long long)c * ldd] = alpha * src[r + (long long)c * lds];
  }
}
int copy_matrix(bk_ctx* ctx, const double* src, long long lds, int rows, int cols, double alpha,
                double* dst, long long ldd) {
  if (rows <= 0 || cols <= 0) return BK_OK;
  copy_matrix_kernel<<<grid_for((long long)rows * cols, 256, 32 * ctx->sm_count), 256, 0,
                       ctx->stream>>>(src, lds, rows, cols, alpha, dst, ldd);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

__global__ void gather_columns_kernel(const double* __restrict__ src, long long lds, int rows,
                                      int cols, const int* __restrict__ perm,
                                      double* __restrict__ dst, long long ldd) {
  const long long total = (long long)rows * cols;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(idx % rows), c = (int)(idx / rows);
    dst[r + (long long)c * ldd] = src[r + (long long)perm[c] * lds];
  }
}
int gather_columns(bk_ctx* ctx, const double* src, long long lds, int rows, int cols,
                   const int* perm, double* dst, long long ldd) {
  if (rows <= 0 || cols <= 0) return BK_OK;
  gather_columns_kernel<<<grid_for((long long)rows * cols, 256, 32 * ctx->sm_count), 256, 0,
                          ctx->stream>>>(src, lds, rows, cols, perm, dst, ldd);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

__global__ void fill_kernel(double* __restrict__ p, long long n, double v) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    p[i] = v;
}
int fill(bk_ctx* ctx, double* p, long long n, double v) {
  if (n <= 0) return BK_OK;
  fill_kernel<<<grid_for(n, 256, 32 * ctx->sm_count), 256, 0, ctx->stream>>>(p, n, v);
  BK_LAUNCHED(ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

}  // namespace bk
