// Context, error plumbing, pinned host memory, micro-benchmarks for the roofline denominators.
#include <cstring>
#include "common.cuh"
#include "dgemm.cuh"
#include "kernels.cuh"

#include <map>
#include <mutex>
#include <tuple>

namespace bk {
namespace {
struct BigCache {
  std::mutex mu;
  std::map<std::tuple<int, cudaStream_t, size_t>, std::vector<void*>> free_blocks;
};
BigCache& big_cache() {
  static BigCache c;
  return c;
}
}  // namespace

void* big_cache_get(size_t bytes, cudaStream_t st, int dev) {
  {
    BigCache& c = big_cache();
    std::lock_guard<std::mutex> lk(c.mu);
    auto it = c.free_blocks.find(std::make_tuple(dev, st, bytes));
    if (it != c.free_blocks.end() && !it->second.empty()) {
      void* p = it->second.back();
      it->second.pop_back();
      return p;
    }
  }
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) {
    // out of memory with blocks parked in the free list: give them back and retry once
    cudaGetLastError();
    big_cache_trim(dev);
    if (cudaMalloc(&p, bytes) != cudaSuccess) return nullptr;
  }
  return p;
}
void big_cache_put(void* p, size_t bytes, cudaStream_t st, int dev) {
  BigCache& c = big_cache();
  std::lock_guard<std::mutex> lk(c.mu);
  c.free_blocks[std::make_tuple(dev, st, bytes)].push_back(p);
}
void big_cache_trim(int dev) {
  BigCache& c = big_cache();
  std::vector<void*> victims;
  {
    std::lock_guard<std::mutex> lk(c.mu);
    for (auto it = c.free_blocks.begin(); it != c.free_blocks.end();) {
      if (std::get<0>(it->first) == dev) {
        victims.insert(victims.end(), it->second.begin(), it->second.end());
        it = c.free_blocks.erase(it);
      } else {
        ++it;
      }
    }
  }
  if (!victims.empty()) cudaDeviceSynchronize();
  for (void* p : victims) cudaFree(p);
}

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
  {
    int lo = 0, hi = 0;
    BK_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    BK_CUDA(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, hi));
  }
  if (bk::pool_enabled()) {
    // keep freed blocks cached in the device's default pool (trimmed again in bk_destroy)
    cudaMemPool_t pool;
    BK_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    BK_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  }
  bk::alloc_stream() = c->stream;
  c->copier = bk::copier_create(device, c->copy_stream);
  c->gemm_ws_side.plain = true;
  c->panel_cache[0].plain = c->panel_cache[1].plain = true;
  c->counters.plain = true;
  BK_TRY(c->gemm_ws_side.alloc((size_t)4 << 20));
  BK_CUDA(cudaMallocHost((void**)&c->host_scratch, 4096));
  BK_TRY(c->counters.alloc(64));
  BK_CUDA(cudaMemsetAsync(c->counters.p, 0, 64 * sizeof(unsigned), c->stream));
  *out = c;
  return BK_OK;
}

void bk_destroy(bk_ctx* ctx) {
  if (!ctx) return;
  bk::bind_ctx(ctx);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamSynchronize(ctx->copy_stream);
  bk::copier_destroy(ctx->copier);
  ctx->copier = nullptr;
  ctx->gemm_ws.release();
  ctx->gemm_ws_side.release();
  ctx->panel_cache[0].release();
  ctx->panel_cache[1].release();
  for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
  ctx->event_pool.clear();
  ctx->barrier.release();
  ctx->scratch.release();
  ctx->counters.release();
  if (ctx->host_scratch) cudaFreeHost(ctx->host_scratch);
  ctx->host_scratch = nullptr;
  for (auto& w : ctx->ws) w.release();
  cudaStreamSynchronize(ctx->stream);
  bk::big_cache_trim(ctx->device);
  if (bk::pool_enabled()) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0);
  }
  if (bk::alloc_stream() == ctx->stream) bk::alloc_stream() = nullptr;
  cudaStreamDestroy(ctx->stream);
  cudaStreamDestroy(ctx->copy_stream);
  cudaStreamDestroy(ctx->side_stream);
  delete ctx;
}

int bk_trim(bk_ctx* ctx) {
  BK_REQUIRE(ctx != nullptr, "bk_trim: ctx is NULL");
  BK_CUDA(bk::bind_ctx(ctx));
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  for (auto& w : ctx->ws) w.release();
  ctx->gemm_ws.release();
  ctx->panel_cache[0].release();
  ctx->panel_cache[1].release();
  BK_CUDA(cudaStreamSynchronize(ctx->stream));
  bk::big_cache_trim(ctx->device);
  if (bk::pool_enabled()) {
    cudaMemPool_t pool;
    BK_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
    BK_CUDA(cudaMemPoolTrimTo(pool, 0));
  }
  return BK_OK;
}

int bk_device_info(bk_ctx* ctx, char* name, int name_len, int* sm_count, int64_t* hbm_total,
                   int64_t* hbm_free) {
  BK_REQUIRE(ctx != nullptr, "bk_device_info: ctx is NULL");
  BK_CUDA(bk::bind_ctx(ctx));
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
  BK_CUDA(bk::bind_ctx(ctx));
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

// both at once: 4 DMMA + 32 DFMA per warp and iteration (equal pipe time if the pipes are separate: 512 vs 64 flops
// per warp instruction).  Tells whether the FP64 tensor path and the FP64 FMA path are the same hardware.
__global__ void __launch_bounds__(256) dmma_dfma_mix_kernel(double* out, int iters) {
  double c[4][2], a[32];
#pragma unroll
  for (int i = 0; i < 4; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 32; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double x = 1.0 + 1e-9 * threadIdx.x, y = 1.0 - 1e-9 * threadIdx.x, bb = 1.0000001, cc = 1e-12;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      dmma884(c[i][0], c[i][1], x, y);
#pragma unroll
      for (int q = 0; q < 8; ++q) a[i * 8 + q] = fma(a[i * 8 + q], bb, cc);
    }
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 4; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 32; ++i) s += a[i];
  if (s == 123.456) out[0] = s;
}

// Latencies of dependent FP64 operations, one warp on one SM, cycles per operation measured with clock64():
// which = 0 DFMA, 1 DMMA.8x8x4 (accumulator chain), 2 DADD, 3 64-bit shuffle (two SHFL) + DADD, 4 shared-memory
// round trip (store, load of the neighbour's value).  The chains of the chasing kernel, the panel QR and the divide &
// conquer are made of these.
__global__ void __launch_bounds__(32) fp64_latency_kernel(double* out, int which, int iters) {
  __shared__ double sm[64];
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0000001, c = 1e-12, c1 = 0.0;
  sm[threadIdx.x] = a;
  __syncwarp();
  const long long t0 = clock64();
  if (which == 0) {
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fma(a, b, c);
  } else if (which == 1) {
#pragma unroll 16
    for (int i = 0; i < iters; ++i) dmma884(a, c1, b, c);
  } else if (which == 2) {
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = a + c;
  } else if (which == 3) {
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = a + __shfl_xor_sync(0xffffffffu, a, 1);
  } else {
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
      sm[threadIdx.x] = a;
      __syncwarp();
      a = sm[threadIdx.x ^ 1] + c;
      __syncwarp();
    }
  }
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[0] = (double)(t1 - t0) / iters;
  if (a + c1 == 123.456) out[1] = a;
}

__global__ void copy_peak_kernel(const double2* __restrict__ src, double2* __restrict__ dst,
                                 long long n2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2;
       i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// read-only streaming ceilings: (a) plain vector loads, (b) TMA bulk copies into an mbarrier ring
__global__ void __launch_bounds__(512) read_peak_kernel(const double2* __restrict__ src, long long n2,
                                                        double* __restrict__ out) {
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n2; i += 4 * stride) {
    const double2 a = src[i], b = src[i + stride], c = src[i + 2 * stride], d = src[i + 3 * stride];
    s0 += a.x + a.y;
    s1 += b.x + b.y;
    s2 += c.x + c.y;
    s3 += d.x + d.y;
  }
  for (; i < n2; i += stride) s0 += src[i].x + src[i].y;
  const double s = s0 + s1 + s2 + s3;
  if (s == 123.456) out[0] = s;
}

__device__ __forceinline__ void mb_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}
// 16 consumer warps + 1 producer warp; each stage = 8 x 4 KB bulk copies (32 KB), 5 stages
__global__ void __launch_bounds__(544, 1) tma_read_peak_kernel(const double* __restrict__ src, long long chunks,
                                                               double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char dsm[];
  double* ring = reinterpret_cast<double*>(dsm);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + 5 * 4096);
  uint64_t* empty = full + 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&empty[i])), "r"(16) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  unsigned it = 0;
  double acc = 0.0;
  if (threadIdx.x == 512) {
    for (long long c = blockIdx.x; c < chunks; c += gridDim.x, ++it) {
      const int st = it % 5;
      mb_wait(&empty[st], ((it / 5) & 1u) ^ 1u);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full[st])), "r"(32768u) : "memory");
      for (int k = 0; k < 8; ++k)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                         smem_u32(ring + st * 4096 + k * 512)),
                     "l"(src + c * 4096 + k * 512), "r"(4096u), "r"(smem_u32(&full[st]))
                     : "memory");
    }
  } else if (threadIdx.x < 512) {
    for (long long c = blockIdx.x; c < chunks; c += gridDim.x, ++it) {
      const int st = it % 5;
      mb_wait(&full[st], (it / 5) & 1u);
      const double* sm = ring + st * 4096 + threadIdx.x;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += sm[k * 512];
      __syncwarp();
      if ((threadIdx.x & 31) == 0)
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[st])) : "memory");
    }
  }
  if (acc == 123.456) out[0] = acc;
}

// strided variant of the TMA ring: the access pattern of the tridiagonalisation SYMV - 8 columns x 512
// rows per stage, S consecutive row chunks per column before moving to the next 8 columns.
__global__ void __launch_bounds__(544, 1) tma_strided_peak_kernel(const double* __restrict__ src, int n, long long lda,
                                                                  int S, double* __restrict__ out) {
  extern __shared__ __align__(128) unsigned char dsm[];
  double* ring = reinterpret_cast<double*>(dsm);
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + 5 * 4096);
  uint64_t* empty = full + 5;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 5; ++i) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&full[i])), "r"(1) : "memory");
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(&empty[i])), "r"(16) : "memory");
    }
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
  __syncthreads();
  const int colblocks = n / 8, rowsegs = n / (512 * S);
  const long long blocks = (long long)colblocks * rowsegs;
  unsigned it = 0;
  double acc = 0.0;
  if (threadIdx.x == 512) {
    for (long long b = blockIdx.x; b < blocks; b += gridDim.x) {
      const long long cbk = b % colblocks, rs = b / colblocks;
      for (int s = 0; s < S; ++s, ++it) {
        const int st = it % 5;
        mb_wait(&empty[st], ((it / 5) & 1u) ^ 1u);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&full[st])), "r"(32768u) : "memory");
        const double* base = src + (rs * S + s) * 512 + cbk * 8 * lda;
        for (int k = 0; k < 8; ++k)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                           smem_u32(ring + st * 4096 + k * 512)),
                       "l"(base + k * lda), "r"(4096u), "r"(smem_u32(&full[st]))
                       : "memory");
      }
    }
  } else if (threadIdx.x < 512) {
    for (long long b = blockIdx.x; b < blocks; b += gridDim.x) {
      for (int s = 0; s < S; ++s, ++it) {
        const int st = it % 5;
        mb_wait(&full[st], (it / 5) & 1u);
        const double* sm = ring + st * 4096 + threadIdx.x;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc += sm[k * 512];
        __syncwarp();
        if ((threadIdx.x & 31) == 0)
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(&empty[st])) : "memory");
      }
    }
  }
  if (acc == 123.456) out[0] = acc;
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
  BK_CUDA(bk::bind_ctx(ctx));
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
  if (kind == 7) {  // size = which (see fp64_latency_kernel); returns cycles per dependent operation
    BK_TRY(buf.alloc(16));
    if (iters <= 0) iters = 4096;
    double h[2] = {0.0, 0.0};
    for (int rep = 0; rep < 2; ++rep) {
      bk::fp64_latency_kernel<<<1, 32, 0, ctx->stream>>>(buf.p, (int)size, iters);
      BK_LAUNCHED(ctx);
      BK_CUDA(cudaGetLastError());
      BK_CUDA(cudaMemcpyAsync(h, buf.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
      BK_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    *result = h[0];
    return BK_OK;
  }
  if (kind == 6) {
    BK_TRY(buf.alloc(16));
    const int blocks = ctx->sm_count * 8;
    if (iters <= 0) iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      tm.start();
      bk::dmma_dfma_mix_kernel<<<blocks, 256, 0, ctx->stream>>>(buf.p, iters);
      BK_LAUNCHED(ctx);
      const double s = tm.stop();
      BK_CUDA(cudaGetLastError());
      const double flops = (double)blocks * iters * (8.0 * 4.0 * 512.0 + 256.0 * 32.0 * 2.0);
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
  if (kind == 3 || kind == 4) {
    const long long n = (size > 0) ? size : (1LL << 29);  // doubles (4 GiB)
    bk::DevBuf<double> o;
    BK_TRY(buf.alloc((size_t)n));
    BK_TRY(o.alloc(16));
    bk::fill_pattern_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(buf.p, n, 1);
    double best = 0.0;
    if (iters <= 0) iters = 10;
    const size_t smem = 5 * 32768 + 128;
    if (kind == 4)
      BK_CUDA(cudaFuncSetAttribute(bk::tma_read_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    for (int rep = 0; rep < iters; ++rep) {
      tm.start();
      if (kind == 3)
        bk::read_peak_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>((const double2*)buf.p, n / 2, o.p);
      else
        bk::tma_read_peak_kernel<<<ctx->sm_count, 544, smem, ctx->stream>>>(buf.p, n / 4096, o.p);
      BK_LAUNCHED(ctx);
      const double s = tm.stop();
      BK_CUDA(cudaGetLastError());
      best = fmax(best, 8.0 * (double)n / s * 1e-9);
    }
    *result = best;
    return BK_OK;
  }
  if (kind == 5) {
    // size = S (row chunks per column run); matrix 16384 x 16384, lda 16400
    const int n = 16384;
    const long long lda = 16400;
    const int S = (size > 0) ? (int)size : 2;
    bk::DevBuf<double> o;
    BK_TRY(buf.alloc((size_t)lda * n));
    BK_TRY(o.alloc(16));
    bk::fill_pattern_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(buf.p, lda * n, 1);
    const size_t smem = 5 * 32768 + 128;
    BK_CUDA(cudaFuncSetAttribute(bk::tma_strided_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    double best = 0.0;
    if (iters <= 0) iters = 5;
    for (int rep = 0; rep < iters; ++rep) {
      tm.start();
      bk::tma_strided_peak_kernel<<<ctx->sm_count, 544, smem, ctx->stream>>>(buf.p, n, lda, S, o.p);
      BK_LAUNCHED(ctx);
      const double sec = tm.stop();
      BK_CUDA(cudaGetLastError());
      best = fmax(best, 8.0 * (double)n * n / sec * 1e-9);
    }
    *result = best;
    return BK_OK;
  }
  bk::set_error("bk_microbench: unknown kind %d", kind);
  return BK_ERR_ARG;
}

int bk_dgemm_bench(bk_ctx* ctx, int ta, int tb, int64_t m, int64_t n, int64_t k, int lower, double beta,
                   int iters, double* seconds) {
  BK_REQUIRE(ctx && seconds && m > 0 && n > 0 && k > 0, "bk_dgemm_bench: bad arguments");
  BK_CUDA(bk::bind_ctx(ctx));
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
  BK_CUDA(cudaMemsetAsync(C.p, 0, sizeof(double) * (size_t)m * n, ctx->stream));
  BK_TRY(bk::gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, 1e-3, A.p, lda, B.p, ldb, beta, C.p, m, lower));
  tm.start();
  for (int i = 0; i < iters; ++i)
    BK_TRY(bk::gemm(ctx, ta != 0, tb != 0, (int)m, (int)n, (int)k, 1e-3, A.p, lda, B.p, ldb, beta, C.p, m, lower));
  *seconds = tm.stop() / iters;
  return BK_OK;
}

}  // extern "C"
