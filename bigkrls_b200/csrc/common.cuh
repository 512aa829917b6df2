// Shared internals of libbigkrls_b200: context, error plumbing, device buffers, small
// device helpers (cp.async, DMMA m8n8k4, grid barrier).  sm_100a only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdarg.h>
#include <time.h>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "../../include/bigkrls_b200.h"

namespace bk {

void set_error(const char* fmt, ...);

#define BK_CUDA(call)                                                                   \
  do {                                                                                  \
    cudaError_t e__ = (call);                                                           \
    if (e__ != cudaSuccess) {                                                           \
      bk::set_error("%s:%d CUDA error %s (%s)", __FILE__, __LINE__, cudaGetErrorName(e__), \
                    cudaGetErrorString(e__));                                           \
      return BK_ERR_CUDA;                                                               \
    }                                                                                   \
  } while (0)

#define BK_TRY(call)                 \
  do {                               \
    int s__ = (call);                \
    if (s__ != BK_OK) return s__;    \
  } while (0)

#define BK_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      bk::set_error(__VA_ARGS__);    \
      return BK_ERR_ARG;             \
    }                                \
  } while (0)

// RAII device buffer (doubles unless stated).  Allocation failures surface as BK_ERR_CUDA.
// Device buffers come from the device's default stream-ordered memory pool (cudaMallocAsync) on the stream of
// the context bound to the calling thread: repeated fits reuse the cached blocks instead of paying the
// map/unmap cost of cudaMalloc/cudaFree for multi-GB matrices.  BK_NO_POOL=1 switches back to cudaMalloc.
inline cudaStream_t& alloc_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
inline bool pool_enabled() {
  static const bool on = (getenv("BK_NO_POOL") == nullptr);
  return on;
}

// Large device blocks (>= 8 MB) bypass the driver's stream-ordered pool: they are cudaMalloc'ed once and parked in
// a per-(device, stream) exact-size free list when released, so that the multi-GB matrices of a fit (K, the
// eigenvector block, the two vcov matrices) are the SAME blocks in every fit.  The driver pool was measured to miss
// on them (60-700 ms per 3.2 GB block, every fit) as soon as smaller allocations interleave with them; its reuse
// is a heuristic, this is not.  Reuse is safe in stream order because a block only ever returns to the stream it
// was used on.  bk_trim / bk_destroy give the parked blocks back.
void* big_cache_get(size_t bytes, cudaStream_t st, int dev);
void big_cache_put(void* p, size_t bytes, cudaStream_t st, int dev);
void big_cache_trim(int dev);
static constexpr size_t kBigBlockBytes = (size_t)8 << 20;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  cudaStream_t pool_stream = nullptr;  // stream the block was allocated on (stream-ordered pool)
  bool pooled = false;
  bool cached = false;  // block belongs to the big-block free list (see big_cache_get)
  int dev = 0;
  bool plain = false;  // long-lived cache: plain cudaMalloc, never from the stream-ordered pool (a persistent block
                       // in the middle of the pool fragments it and later multi-GB allocations fall off the fast path)
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p && borrowed) {
      p = nullptr;
      n = 0;
      borrowed = false;
      return;
    }
    if (p && cached) {
      big_cache_put(p, n * sizeof(T), pool_stream, dev);
    } else if (p) {
      // stream-ordered free: the block returns to the pool once the work queued so far has used it
      if (!pooled || cudaFreeAsync(p, pool_stream) != cudaSuccess) {
        cudaGetLastError();
        cudaFree(p);
      }
    }
    p = nullptr;
    n = 0;
    pooled = false;
    cached = false;
  }
  int alloc(size_t count) {
    release();
    if (count == 0) count = 1;
    cudaError_t e;
    static const bool verbose = getenv("BK_ALLOC_VERBOSE") != nullptr;
    struct timespec t0_, t1_;
    if (verbose) clock_gettime(CLOCK_MONOTONIC, &t0_);
    if (pool_enabled() && !plain && count * sizeof(T) >= kBigBlockBytes) {
      pool_stream = alloc_stream();
      cudaGetDevice(&dev);
      p = (T*)big_cache_get(count * sizeof(T), pool_stream, dev);
      cached = (p != nullptr);
      e = cached ? cudaSuccess : cudaGetLastError();
      if (!cached && e == cudaSuccess) e = cudaErrorMemoryAllocation;
    } else if (pool_enabled() && !plain) {
      pool_stream = alloc_stream();
      e = cudaMallocAsync((void**)&p, count * sizeof(T), pool_stream);
      pooled = (e == cudaSuccess);
    } else {
      e = cudaMalloc((void**)&p, count * sizeof(T));
    }
    if (verbose) {
      clock_gettime(CLOCK_MONOTONIC, &t1_);
      const double ms = (t1_.tv_sec - t0_.tv_sec) * 1e3 + (t1_.tv_nsec - t0_.tv_nsec) * 1e-6;
      if (ms > 1.0) fprintf(stderr, "[alloc] %.1f MB took %.2f ms (%s)\n", count * sizeof(T) * 1e-6, ms, (pooled ? "pool" : "cudaMalloc"));
    }
    if (e != cudaSuccess) {
      p = nullptr;
      pooled = false;
      cudaGetLastError();
      set_error("device allocation of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
      return BK_ERR_CUDA;
    }
    n = count;
    return BK_OK;
  }
  // grow-only
  int ensure(size_t count) {
    if (count <= n && p) return BK_OK;
    return alloc(count);
  }
  // Use (a prefix of) a long-lived cache buffer instead of an allocation of one's own: the cache grows if it has
  // to and is never freed by this object.  For the multi-GB transient work matrices of the eigensolver, whose
  // allocate/free churn fragments the pool and occasionally stalls a fit for hundreds of milliseconds.
  int borrow(DevBuf& cache, size_t count) {
    release();
    int rc = cache.ensure(count);
    if (rc != BK_OK) return rc;
    p = cache.p;
    n = count;
    borrowed = true;
    return BK_OK;
  }
  bool borrowed = false;
  // exchange the blocks (and their provenance) of two buffers
  void swap_block(DevBuf& o) {
    std::swap(p, o.p);
    std::swap(n, o.n);
    std::swap(pool_stream, o.pool_stream);
    std::swap(pooled, o.pooled);
    std::swap(cached, o.cached);
    std::swap(dev, o.dev);
    std::swap(plain, o.plain);
    std::swap(borrowed, o.borrowed);
  }
};

struct Timer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t s = nullptr;
  int init(cudaStream_t st) {
    s = st;
    BK_CUDA(cudaEventCreate(&a));
    BK_CUDA(cudaEventCreate(&b));
    return BK_OK;
  }
  ~Timer() {
    if (a) cudaEventDestroy(a);
    if (b) cudaEventDestroy(b);
  }
  void start() { cudaEventRecord(a, s); }
  // returns seconds; synchronises the stream
  double stop() {
    cudaEventRecord(b, s);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return (double)ms * 1e-3;
  }
};

// device -> host delivery into caller-owned (possibly pageable) memory, hostcopy.cu
struct HostCopier;
struct CopyTicket {
  std::shared_ptr<void> job;
};

}  // namespace bk

struct bk_ctx {
  int device = 0;
  int sm_count = 0;
  int max_smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t side_stream = nullptr;  // high-priority stream for look-ahead work (panel factorisation under an update)
  bk::DevBuf<double> gemm_ws;          // split-K partials
  bk::DevBuf<double> gemm_ws_side;     // the same for GEMMs issued on the side stream
  bk::DevBuf<double> panel_cache[2];   // [V W] / [W V] panel buffers of the dense->band stage (grow-only)
  std::vector<cudaEvent_t> event_pool; // timing / ordering events reused across fits (created on first use)
  bk::DevBuf<unsigned int> barrier;    // grid-barrier counters
  bk::DevBuf<unsigned char> scratch;   // small general scratch (descriptors, partial sums)
  bk::DevBuf<unsigned int> counters;   // 64 arrival counters, zero between kernels (last-CTA-done reductions)
  bk::DevBuf<double> ws[7];            // cached N x N work matrices of the eigensolver (0 work copy of K,
                                       // 1 stage-2 reflectors, 2-4 divide & conquer, 5-6 its factored top level),
                                       // released by bk_trim / bk_destroy
  double* host_scratch = nullptr;      // 4 KB of pinned host memory for small result read-backs (LOO losses per pass)
  uint64_t n_launches = 0;             // kernels launched through this context (bench "gpu_launches")
  bk::HostCopier* copier = nullptr;    // hostcopy.cu
};

namespace bk {

// makes ctx the current context of the calling thread: device + allocation stream
inline cudaError_t bind_ctx(bk_ctx* c) {
  alloc_stream() = c->stream;
  return cudaSetDevice(c->device);
}

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned smem_u32(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}

// 8-byte async global->shared copy; src_bytes in {0, 8}: 0 zero-fills.
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
// 16-byte async copy; src_bytes in {0, 8, 16}; remaining bytes are zero-filled.
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem)), "l"(gmem),
               "r"(src_bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// FP64 tensor-core MMA, D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds
//   A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4) + {0,1}].   SASS: DMMA.8x8x4
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum; `red` must hold >= 32 doubles of shared memory.  All threads get the result.
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  double t = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0;
  if (wid == 0) {
    t = warp_sum(t);
    if (lane == 0) red[0] = t;
  }
  __syncthreads();
  t = red[0];
  return t;
}

// ---- mbarrier / TMA bulk-copy primitives (sm_90+ PTX; SASS: SYNCS.*, UBLKCP) -----------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  unsigned ok;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// contiguous global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar`
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes,
                                             uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// orders generic-proxy accesses (already performed or acquired) before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;\n" ::: "memory"); }

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier for persistent kernels launched with cudaLaunchCooperativeKernel (all
// CTAs co-resident).  `counter` is monotonically increasing and must be 0 at kernel start;
// `epoch` is a per-thread register counting barriers passed so far.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  epoch += 1;
  if (threadIdx.x == 0) {
    // release-add / acquire-load: cumulative over the CTA barrier above, no full (sc) fences needed
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;\n" ::"l"(counter) : "memory");
    const unsigned target = epoch * gridDim.x;
    while (ld_acquire_u32(counter) < target) {
    }
  }
  __syncthreads();
}

#endif  // __CUDACC__

HostCopier* copier_create(int device, cudaStream_t copy_stream);
void copier_destroy(HostCopier* c);
int copier_submit(bk_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t after, CopyTicket* out);
int copier_wait(CopyTicket* t);
int copy_to_host(bk_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t after);

// Work issued inside the scope goes to the side stream (with its own split-K workspace) instead of the main one.
struct SideStreamScope {
  bk_ctx* c;
  cudaStream_t saved;
  SideStreamScope(bk_ctx* ctx) : c(ctx), saved(ctx->stream) {
    c->stream = c->side_stream;
    c->gemm_ws.swap_block(c->gemm_ws_side);
  }
  ~SideStreamScope() {
    c->stream = saved;
    c->gemm_ws.swap_block(c->gemm_ws_side);
  }
};
// i-th event of the context's pool (timing enabled), created on first use
inline cudaEvent_t pool_event(bk_ctx* c, size_t i) {
  while (c->event_pool.size() <= i) {
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    c->event_pool.push_back(e);
  }
  return c->event_pool[i];
}

// launch-count bookkeeping
#define BK_LAUNCHED(ctx) ((ctx)->n_launches++)

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace bk
