// Context, error plumbing, pinned host memory, micro-benchmarks for the roofline denominators.
#include <cstring>
#include "common.cuh"
#include "dgemm.cuh"
#include "kernels.cuh"

namespace bk {
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace bk

extern "C" {

int bk_version(void) { return BK_VERSION; }
const char* bk_last_error(void) { return bk::g_err; }

int bk_init(int device, bk_ctx** out) {
  BK_REQUIRE(out != nullptr, "bk_init: out is NULL");
  *out = nullptr;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    bk::set_error("bk_init: no CUDA device available (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return BK_ERR_CUDA;
  }
  BK_REQUIRE(device >= 0 && device < count, "bk_init: device %d out of range (0..%d)", device,
             count - 1);
  BK_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  BK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    bk::set_error("bk_init: device %s is sm_%d%d; this library is built for sm_100a only", prop.name,
                  prop.major, prop.minor);
    return BK_ERR_CUDA;
  }
  bk_ctx* c = new bk_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  c->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
  BK_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  BK_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  *out = c;
  return BK_OK;
}

void bk_destroy(bk_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  ctx->gemm_ws.release();
  ctx->barrier.release();
  ctx->scratch.release();
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  delete ctx;
}

int bk_device_info(bk_ctx* ctx, char* name, int name_len, int* sm_count, int64_t* hbm_total,
                   int64_t* hbm_free) {
  BK_REQUIRE(ctx != nullptr, "bk_device_info: ctx is NULL");
  BK_CUDA(cudaSetDevice(ctx->device));
  cudaDeviceProp prop;
  BK_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
  if (name && name_len > 0) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  size_t f = 0, t = 0;
  BK_CUDA(cudaMemGetInfo(&f, &t));
  if (hbm_total) *hbm_total = (int64_t)t;
  if (hbm_free) *hbm_free = (int64_t)f;
  return BK_OK;
}

int64_t bk_launch_count(bk_ctx* ctx) { return ctx ? (int64_t)ctx->n_launches : 0; }

int bk_host_alloc(bk_ctx* ctx, int64_t bytes, void** out) {
  BK_REQUIRE(ctx && out && bytes > 0, "bk_host_alloc: bad arguments");
  BK_CUDA(cudaSetDevice(ctx->device));
  BK_CUDA(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault));
  return BK_OK;
}
int bk_host_free(bk_ctx* ctx, void* p) {
  BK_REQUIRE(ctx != nullptr, "bk_host_free: ctx is NULL");
  if (p) BK_CUDA(cudaFreeHost(p));
  return BK_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------
// micro-benchmarks (roofline denominators for FP64, which MEASURED_PEAKS.json does not hold)
// ---------------------------------------------------------------------------------------------
namespace bk {

__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters) {
  double a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(256) dmma_peak_kernel(double* out, int iters) {
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  const double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) dmma884(c[i][0], c[i][1], a, b);
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  if (s == 123.456) out[0] = s;
}

__global__ void copy_peak_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                 long long n2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

__global__ void fill_pattern_kernel(double* p, long long n, unsigned long long seed) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    unsigned long long x = (unsigned long long)i * 6364136223846793005ULL + seed;
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    p[i] = (double)(x >> 11) * (1.0 / 9007199254740992.0) - 0.5;
  }
}

}  // namespace bk

extern "C" {

int bk_microbench(bk_ctx* ctx, int kind, int64_t size, int iters, double* result) {
  BK_REQUIRE(ctx && result, "bk_microbench: bad arguments");
  BK_CUDA(cudaSetDevice(ctx->device));
  bk::Timer tm;
  BK_TRY(tm.init(ctx->stream));
  bk::DevBuf<double> buf;
  if (kind == 0 || kind == 1) {
    BK_TRY(buf.alloc(16));
    const int blocks = ctx->sm_count * 8;
    if (iters <= 0) iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {  // rep 0 = warm-up
      tm.start();
      if (kind == 0)
        bk::dfma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(buf.p, iters);
      else
        bk::dmma_peak_kernel<<<blocks, 256, 0, ctx->stream>>>(buf.p, iters);
      BK_LAUNCHED(ctx);
      const double s = tm.stop();
      BK_CUDA(cudaGetLastError());
      // DFMA: 2 flops per lane-op; DMMA m8n8k4: 8*8*4*2 flops per warp instruction
      const double flops = (kind == 0) ? (double)blocks * 256.0 * 16.0 * iters * 2.0
                                       : (double)blocks * 8.0 * 16.0 * iters * 512.0;
      *result = flops / s * 1e-12;
    }
    return BK_OK;
  }
  if (kind == 2) {
    const long long n = (size > 0) ? size : (1LL << 28);  // doubles (2 GiB)
    bk::DevBuf<double> dst;
    BK_TRY(buf.alloc((size_t)n));
    BK_TRY(dst.alloc((size_t)n));
    bk::fill_pattern_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(buf.p, n, 1);
    double best = 0.0;
    if (iters <= 0) iters = 10;
    for (int rep = 0; rep < iters; ++rep) {
      tm.start();
      bk::copy_peak_kernel<<<ctx->sm_count * 16, 256, 0, ctx->stream>>>(
          (const double2*)buf.p, (double2*)dst.p, n / 2);
      BK_LAUNCHED(ctx);
      const double s = tm.stop();
      BK_CUDA(cudaGetLastError());
      best = fmax(best, 16.0 * (double)n / s * 1e-9);
    }
    *result = best;
    return BK_OK;
  }
  bk::set_error("bk_microbench: unknown kind %d", kind);
  return BK_ERR_ARG;
}

int bk_dgemm_bench(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, int lower,
                   int iters, double* seconds) {
  BK_REQUIRE(ctx && seconds && m > 0 && n > 0 && k > 0, "bk_dgemm_bench: bad arguments");
  BK_CUDA(cudaSetDevice(ctx->device));
  bk::DevBuf<double> A, B, C;
  BK_TRY(A.alloc((size_t)m * k));
  BK_TRY(B.alloc((size_t)k * n));
  BK_TRY(C.alloc((size_t)m * n));
  bk::fill_pattern_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(A.p, m * k, 11);
  bk::fill_pattern_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(B.p, k * n, 22);
  const long long lda = ta ? k : m, ldb = tb ? n : k;
  bk::Timer tm;
  BK_TRY(tm.init(ctx->stream));
  if (iters <= 0) iters = 3;
  BK_TRY(bk::gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, 1.0, A.p, lda, B.p, ldb, 0.0, C.p, m,
                  lower != 0));
  tm.start();
  for (int i = 0; i < iters; ++i)
    BK_TRY(bk::gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, 1.0, A.p, lda, B.p, ldb, 0.0, C.p,
                    m, lower != 0));
  *seconds = tm.stop() / iters;
  return BK_OK;
}

}  // extern "C"
