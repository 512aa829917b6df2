// Two-process CUDA IPC probe (one process per GPU, like torchrun ranks): does cudaIpcOpenMemHandle work in this
// container, what do peer stores over NVLink deliver, and what does a flag round trip between two GPUs cost?
// These numbers size the peer-memory collectives of bigkrls_b200/csrc/peer.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ipc_probe tools/ipc_probe.cu && build/ipc_probe
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/wait.h>
#include <unistd.h>

#define CK(x)                                                                          \
  do {                                                                                 \
    cudaError_t e = (x);                                                               \
    if (e != cudaSuccess) {                                                            \
      fprintf(stderr, "[rank %d] %s:%d %s\n", g_rank, __FILE__, __LINE__, cudaGetErrorString(e)); \
      exit(2);                                                                         \
    }                                                                                  \
  } while (0)
static int g_rank = -1;

__device__ __forceinline__ unsigned ld_acq_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_rel_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__global__ void push_kernel(const double2* __restrict__ src, double2* __restrict__ dst, long long n2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
__global__ void pull_kernel(const double2* __restrict__ src, double2* __restrict__ dst, long long n2) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
    dst[i] = src[i];
}
// ping-pong: rank 0 writes flag i to the peer, waits for the echo in its own memory; rank 1 echoes
__global__ void pingpong_kernel(unsigned* my_flag, unsigned* peer_flag, int iters, int initiator, long long* cycles) {
  const long long t0 = clock64();
  for (int i = 1; i <= iters; ++i) {
    if (initiator) {
      st_rel_sys(peer_flag, (unsigned)i);
      while (ld_acq_sys(my_flag) < (unsigned)i) {
      }
    } else {
      while (ld_acq_sys(my_flag) < (unsigned)i) {
      }
      st_rel_sys(peer_flag, (unsigned)i);
    }
  }
  *cycles = clock64() - t0;
}

static void xchg(int wfd, int rfd, const void* send, void* recv, size_t bytes) {
  if (write(wfd, send, bytes) != (ssize_t)bytes) exit(3);
  size_t got = 0;
  while (got < bytes) {
    ssize_t r = read(rfd, (char*)recv + got, bytes - got);
    if (r <= 0) exit(3);
    got += r;
  }
}

int main() {
  int p01[2], p10[2];
  if (pipe(p01) || pipe(p10)) return 1;
  pid_t pid = fork();
  g_rank = pid == 0 ? 1 : 0;
  const int wfd = g_rank == 0 ? p01[1] : p10[1], rfd = g_rank == 0 ? p10[0] : p01[0];
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < 2) {
    if (g_rank == 0) printf("{\"ipc_probe\": \"needs 2 GPUs, found %d\"}\n", ndev);
    return 0;
  }
  CK(cudaSetDevice(g_rank));
  const size_t bytes = 1ull << 30;
  char* mine = nullptr;
  CK(cudaMalloc(&mine, bytes));
  CK(cudaMemset(mine, 0, bytes));
  cudaIpcMemHandle_t hm, hp;
  CK(cudaIpcGetMemHandle(&hm, mine));
  xchg(wfd, rfd, &hm, &hp, sizeof(hm));
  char* peer = nullptr;
  CK(cudaIpcOpenMemHandle((void**)&peer, hp, cudaIpcMemLazyEnablePeerAccess));
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, g_rank, 1 - g_rank));
  char tok = 1, tok2;
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  double push_gbs[3] = {0, 0, 0}, pull_gbs = 0, memcpy_gbs = 0;
  const size_t sizes[3] = {1u << 20, 16u << 20, 512u << 20};
  // flags live in the first 4 KB; data after 1 MB
  for (int s = 0; s < 3; ++s) {
    const long long n2 = sizes[s] / 16;
    for (int rep = 0; rep < 3; ++rep) {
      xchg(wfd, rfd, &tok, &tok2, 1);
      CK(cudaEventRecord(a));
      push_kernel<<<148 * 4, 256>>>((const double2*)(mine + (1 << 20)), (double2*)(peer + (1 << 20)), n2);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      push_gbs[s] = sizes[s] / (ms * 1e-3) * 1e-9;
    }
  }
  {
    const long long n2 = (512u << 20) / 16;
    for (int rep = 0; rep < 3; ++rep) {
      xchg(wfd, rfd, &tok, &tok2, 1);
      CK(cudaEventRecord(a));
      pull_kernel<<<148 * 4, 256>>>((const double2*)(peer + (1 << 20)), (double2*)(mine + (1 << 20)), n2);
      CK(cudaEventRecord(b));
      CK(cudaEventSynchronize(b));
      float ms;
      CK(cudaEventElapsedTime(&ms, a, b));
      pull_gbs = (512u << 20) / (ms * 1e-3) * 1e-9;
    }
    xchg(wfd, rfd, &tok, &tok2, 1);
    CK(cudaEventRecord(a));
    CK(cudaMemcpyAsync(peer + (1 << 20), mine + (1 << 20), 512u << 20, cudaMemcpyDefault));
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    memcpy_gbs = (512u << 20) / (ms * 1e-3) * 1e-9;
  }
  // flag round trip
  long long* cyc;
  CK(cudaMalloc(&cyc, 8));
  CK(cudaMemset(mine, 0, 4096));
  CK(cudaDeviceSynchronize());
  xchg(wfd, rfd, &tok, &tok2, 1);
  const int iters = 2000;
  pingpong_kernel<<<1, 1>>>((unsigned*)mine, (unsigned*)peer, iters, g_rank == 0, cyc);
  CK(cudaDeviceSynchronize());
  long long hc = 0;
  CK(cudaMemcpy(&hc, cyc, 8, cudaMemcpyDeviceToHost));
  int clk = 0;
  CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, g_rank));
  const double rt_us = (double)hc / iters / (clk * 1e-3);
  xchg(wfd, rfd, &tok, &tok2, 1);
  if (g_rank == 0)
    printf("{\"ipc_probe\": \"ok\", \"can_access_peer\": %d, \"push_GBs_1MB\": %.1f, \"push_GBs_16MB\": %.1f, \"push_GBs_512MB\": %.1f, "
           "\"pull_GBs_512MB\": %.1f, \"memcpy_peer_GBs_512MB\": %.1f, \"flag_round_trip_us\": %.2f}\n",
           can, push_gbs[0], push_gbs[1], push_gbs[2], pull_gbs, memcpy_gbs, rt_us);
  CK(cudaIpcCloseMemHandle(peer));
  CK(cudaFree(mine));
  if (g_rank == 0) {
    int st = 0;
    waitpid(pid, &st, 0);
  }
  return 0;
}
