// Peer-memory communicator (see peer.cuh): symmetric heaps over CUDA IPC, flag-synchronised collectives as
// kernels on the library stream.
#include <cstring>
#include "peer.cuh"

namespace bk {

namespace {

__global__ void peer_barrier_kernel(PeerDev pd, int ch, unsigned seq) {
  const int r = threadIdx.x;
  if (r < pd.world) {
    st_release_sys_u32(peer_flag(pd, r, ch, pd.rank), seq);
    peer_wait_flag(pd, ch, r, seq);
  }
}

__global__ void peer_wait_kernel(PeerDev pd, int ch, unsigned src_mask, unsigned seq) {
  const int r = threadIdx.x;
  if (r < pd.world && (src_mask & (1u << r))) peer_wait_flag(pd, ch, r, seq);
}

// single CTA: stage own vector at every rank, signal, wait for everyone, add the slots in rank order
__global__ void __launch_bounds__(1024) peer_allreduce_kernel(PeerDev pd, double* buf, int n, int parity, unsigned seq) {
  const size_t slot = (size_t)PEER_AR_MAX * sizeof(double);
  const size_t base = PEER_AR_OFF + (size_t)parity * BK_MAX_PEERS * slot;
  for (int r = 0; r < pd.world; ++r) {
    double* dst = reinterpret_cast<double*>(pd.heap[r] + base + (size_t)pd.rank * slot);
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = buf[i];
  }
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < pd.world) {
    st_release_sys_u32(peer_flag(pd, threadIdx.x, CH_COLL, pd.rank), seq);
    peer_wait_flag(pd, CH_COLL, threadIdx.x, seq);
  }
  __syncthreads();
  const char* mine = pd.heap[pd.rank] + base;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < pd.world; ++r) s += __ldcg(reinterpret_cast<const double*>(mine + (size_t)r * slot) + i);
    buf[i] = s;
  }
}

// rows x cols block (column-major, lds) -> heap[dst] + dst_off (ldd) for every dst in the mask, then the flag
__global__ void __launch_bounds__(256) peer_push2d_kernel(PeerDev pd, const double* __restrict__ src, long long lds,
                                                          int rows, int cols, size_t dst_off, long long ldd,
                                                          unsigned dst_mask, int ch, unsigned seq, int cnt_idx) {
  const long long total = (long long)rows * cols;
  const bool vec = ((rows & 1) == 0) && ((lds & 1) == 0) && ((ldd & 1) == 0) && ((((uintptr_t)src) & 15u) == 0) &&
                   ((dst_off & 15u) == 0);
  for (int r = 0; r < pd.world; ++r) {
    if (!(dst_mask & (1u << r))) continue;
    double* dst = reinterpret_cast<double*>(pd.heap[r] + dst_off);
    if (vec) {
      const int rp = rows / 2;
      for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < (long long)rp * cols;
           idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rp) * 2, j = (int)(idx / rp);
        *reinterpret_cast<double2*>(dst + i + (long long)j * ldd) =
            *reinterpret_cast<const double2*>(src + i + (long long)j * lds);
      }
    } else {
      for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
           idx += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(idx % rows), j = (int)(idx / rows);
        dst[i + (long long)j * ldd] = src[i + (long long)j * lds];
      }
    }
  }
  peer_signal_last_cta(pd, peer_counter(pd, cnt_idx), gridDim.x, dst_mask, ch, seq);
}

// n contiguous doubles -> heap[dst] + dst_off for every dst in the mask, then the flag
__global__ void __launch_bounds__(256) peer_push1d_kernel(PeerDev pd, const double* __restrict__ src, long long n,
                                                          size_t dst_off, unsigned dst_mask, int ch, unsigned seq,
                                                          int cnt_idx) {
  const bool vec = ((((uintptr_t)src) & 15u) == 0) && ((dst_off & 15u) == 0);
  const long long n2 = vec ? n / 2 : 0;
  for (int r = 0; r < pd.world; ++r) {
    if (!(dst_mask & (1u << r))) continue;
    double* dst = reinterpret_cast<double*>(pd.heap[r] + dst_off);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n2; i += (long long)gridDim.x * blockDim.x)
      reinterpret_cast<double2*>(dst)[i] = reinterpret_cast<const double2*>(src)[i];
    for (long long i = 2 * n2 + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
      dst[i] = src[i];
  }
  peer_signal_last_cta(pd, peer_counter(pd, cnt_idx), gridDim.x, dst_mask, ch, seq);
}

int push1d(bk_peer* p, const double* src, long long n, size_t dst_off, unsigned dst_mask, int ch, unsigned seq,
           cudaStream_t st) {
  const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(n, 1024), 4LL * p->ctx->sm_count));
  peer_push1d_kernel<<<blocks, 256, 0, st>>>(p->dev, src, n, dst_off, dst_mask, ch, seq, ch);
  BK_LAUNCHED(p->ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

__global__ void __launch_bounds__(256) peer_sum_slots_kernel(PeerDev pd, double* __restrict__ buf, long long n, size_t base,
                                                             long long slot_elems) {
  const double* mine = reinterpret_cast<const double*>(pd.heap[pd.rank] + base);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < pd.world; ++r) s += __ldcg(mine + (long long)r * slot_elems + i);
    buf[i] = s;
  }
}

int open_heaps(bk_peer* p, size_t bytes) {
  // own heap + IPC handle exchange (host-side collective through the bootstrap callback)
  char* mine = nullptr;
  BK_CUDA(cudaMalloc((void**)&mine, bytes));
  BK_CUDA(cudaMemset(mine, 0, PEER_CTRL_BYTES));
  cudaIpcMemHandle_t h;
  BK_CUDA(cudaIpcGetMemHandle(&h, mine));
  std::vector<cudaIpcMemHandle_t> all(p->world);
  if (p->exchange(p->user, &h, all.data(), (int64_t)sizeof(h)) != 0) {
    cudaFree(mine);
    set_error("bk_peer: the bootstrap exchange callback failed");
    return BK_ERR_COMM;
  }
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) {
      p->dev.heap[r] = mine;
      continue;
    }
    void* q = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("bk_peer: cudaIpcOpenMemHandle(rank %d) failed: %s - peer access over NVLink is required", r,
                cudaGetErrorString(e));
      cudaGetLastError();
      return BK_ERR_COMM;
    }
    p->dev.heap[r] = (char*)q;
  }
  p->heap_bytes = bytes;
  p->bump = PEER_CTRL_BYTES;
  // nobody may touch a heap before everyone has mapped it (and zeroed its control area)
  char tok = 1;
  std::vector<char> toks(p->world);
  if (p->exchange(p->user, &tok, toks.data(), 1) != 0) {
    set_error("bk_peer: the bootstrap exchange callback failed");
    return BK_ERR_COMM;
  }
  return BK_OK;
}

int close_heaps(bk_peer* p) {
  if (!p->heap_bytes) return BK_OK;
  BK_CUDA(cudaDeviceSynchronize());
  // everyone is done with everyone's heap before anything is unmapped
  char tok = 1;
  std::vector<char> toks(p->world);
  p->exchange(p->user, &tok, toks.data(), 1);
  for (int r = 0; r < p->world; ++r) {
    if (r != p->rank && p->dev.heap[r]) cudaIpcCloseMemHandle(p->dev.heap[r]);
  }
  p->exchange(p->user, &tok, toks.data(), 1);
  if (p->dev.heap[p->rank]) cudaFree(p->dev.heap[p->rank]);
  for (int r = 0; r < BK_MAX_PEERS; ++r) p->dev.heap[r] = nullptr;
  p->heap_bytes = 0;
  return BK_OK;
}

}  // namespace

int peer_ensure_heap(bk_peer* p, size_t bytes) {
  bytes = (bytes + PEER_CTRL_BYTES + (size_t)(1 << 21) - 1) & ~((size_t)(1 << 21) - 1);
  if (bytes <= p->heap_bytes) {
    peer_reset(p);
    return BK_OK;
  }
  BK_TRY(close_heaps(p));
  // the flags restart from zero with the new heaps
  memset(p->seq, 0, sizeof(p->seq));
  p->ar_count = 0;
  p->arl_count = 0;
  return open_heaps(p, bytes);
}

int peer_alloc(bk_peer* p, size_t bytes, size_t* offset) {
  const size_t a = (p->bump + 255) & ~(size_t)255;
  if (a + bytes > p->heap_bytes) {
    set_error("bk_peer: symmetric heap exhausted (%zu + %zu > %zu bytes)", a, bytes, p->heap_bytes);
    return BK_ERR_COMM;
  }
  *offset = a;
  p->bump = a + bytes;
  return BK_OK;
}

int peer_barrier(bk_peer* p, cudaStream_t st) {
  const unsigned seq = peer_next_seq(p, CH_BARRIER);
  peer_barrier_kernel<<<1, 32, 0, st>>>(p->dev, CH_BARRIER, seq);
  BK_LAUNCHED(p->ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int peer_wait(bk_peer* p, int ch, unsigned src_mask, unsigned seq, cudaStream_t st) {
  peer_wait_kernel<<<1, 32, 0, st>>>(p->dev, ch, src_mask, seq);
  BK_LAUNCHED(p->ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int peer_allreduce_sum(bk_peer* p, double* buf, long long n, cudaStream_t st) {
  for (long long o = 0; o < n; o += PEER_AR_MAX) {
    const int cnt = (int)std::min<long long>(PEER_AR_MAX, n - o);
    const unsigned seq = peer_next_seq(p, CH_COLL);
    peer_allreduce_kernel<<<1, 1024, 0, st>>>(p->dev, buf + o, cnt, (int)(p->ar_count++ & 1u), seq);
    BK_LAUNCHED(p->ctx);
  }
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int peer_allreduce_sum_large(bk_peer* p, double* buf, long long n, size_t stage_off, long long slot_elems, cudaStream_t st) {
  if (n <= 0) return BK_OK;
  const unsigned all = (1u << p->world) - 1u;
  const int q = (int)(p->arl_count++ & 1u);
  const size_t base = stage_off + sizeof(double) * (size_t)q * (size_t)p->world * (size_t)slot_elems;
  const unsigned seq = peer_next_seq(p, CH_ARLARGE);
  BK_TRY(push1d(p, buf, n, base + sizeof(double) * (size_t)p->rank * (size_t)slot_elems, all, CH_ARLARGE, seq, st));
  BK_TRY(peer_wait(p, CH_ARLARGE, all, seq, st));
  const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(n, 512), 4LL * p->ctx->sm_count));
  peer_sum_slots_kernel<<<blocks, 256, 0, st>>>(p->dev, buf, n, base, slot_elems);
  BK_LAUNCHED(p->ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int peer_push2d(bk_peer* p, const double* src, long long lds, int rows, int cols, size_t dst_off, long long ldd,
                unsigned dst_mask, int ch, unsigned seq, cudaStream_t st) {
  const long long total = (long long)rows * cols;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(ceil_div(total, 512), 2LL * p->ctx->sm_count));
  peer_push2d_kernel<<<blocks, 256, 0, st>>>(p->dev, src, lds, rows, cols, dst_off, ldd, dst_mask, ch, seq, ch);
  BK_LAUNCHED(p->ctx);
  BK_CUDA(cudaGetLastError());
  return BK_OK;
}

int peer_allgatherv_sym(bk_peer* p, size_t off, const long long* counts, const long long* displs, cudaStream_t st) {
  const int me = p->rank;
  const unsigned others = ((1u << p->world) - 1u) & ~(1u << me);
  const unsigned seq = peer_next_seq(p, CH_GATHER);
  // my segment -> the same place in every other heap (a zero-length segment still signals)
  BK_TRY(push1d(p, peer_ptr(p, off) + displs[me], counts[me], off + sizeof(double) * (size_t)displs[me], others,
                CH_GATHER, seq, st));
  return peer_wait(p, CH_GATHER, others, seq, st);
}

int peer_broadcast_sym(bk_peer* p, size_t off, long long n, int root, cudaStream_t st) {
  const unsigned others = ((1u << p->world) - 1u) & ~(1u << root);
  const unsigned seq = peer_next_seq(p, CH_BCAST);  // every rank advances the channel
  if (p->rank == root) return push1d(p, peer_ptr(p, off), n, off, others, CH_BCAST, seq, st);
  return peer_wait(p, CH_BCAST, 1u << root, seq, st);
}

int peer_check(bk_peer* p, cudaStream_t st) {
  unsigned e = 0;
  BK_CUDA(cudaMemcpyAsync(&e, p->dev.heap[p->rank] + PEER_ERR_OFF, sizeof(e), cudaMemcpyDeviceToHost, st));
  BK_CUDA(cudaStreamSynchronize(st));
  if (e != 0) {
    set_error("bk_peer: rank %d timed out waiting for rank %u (peer process gone, or ranks out of step)", p->rank,
              e & 0xffu);
    return BK_ERR_COMM;
  }
  return BK_OK;
}

}  // namespace bk

using namespace bk;

extern "C" {

int bk_peer_create(bk_ctx* ctx, int rank, int world, bk_exchange_fn exchange, void* user, bk_peer** out) {
  BK_REQUIRE(ctx && out && exchange, "bk_peer_create: NULL argument");
  BK_REQUIRE(world >= 1 && world <= BK_MAX_PEERS && rank >= 0 && rank < world,
             "bk_peer_create: rank/world out of range (at most %d ranks, one node)", BK_MAX_PEERS);
  BK_CUDA(bind_ctx(ctx));
  bk_peer* p = new bk_peer();
  p->ctx = ctx;
  p->rank = rank;
  p->world = world;
  p->exchange = exchange;
  p->user = user;
  p->dev.rank = rank;
  p->dev.world = world;
  for (int r = 0; r < BK_MAX_PEERS; ++r) p->dev.heap[r] = nullptr;
  const int rc = peer_ensure_heap(p, 64u << 20);
  if (rc != BK_OK) {
    delete p;
    return rc;
  }
  *out = p;
  return BK_OK;
}

void bk_peer_destroy(bk_peer* p) {
  if (!p) return;
  bind_ctx(p->ctx);
  close_heaps(p);
  delete p;
}

// Self-test of the collectives (tests/dist_worker.py, 2+ GPUs): returns the number of mismatches in *failures.
int bk_peer_selftest(bk_peer* p, int* failures) {
  BK_REQUIRE(p && failures, "bk_peer_selftest: NULL argument");
  bk_ctx* ctx = p->ctx;
  BK_CUDA(bind_ctx(ctx));
  cudaStream_t st = ctx->stream;
  const int W = p->world, me = p->rank;
  int bad = 0;
  BK_TRY(peer_ensure_heap(p, 256u << 20));
  // all-reduce, sizes around the chunk limit, repeated (staging parity reuse)
  for (int rep = 0; rep < 3; ++rep)
    for (long long n : {1LL, 16LL, 2860LL, (long long)PEER_AR_MAX, (long long)PEER_AR_MAX + 7, 100003LL}) {
      std::vector<double> h(n);
      for (long long i = 0; i < n; ++i) h[i] = (double)(me + 1) * (double)((i % 97) + rep);
      DevBuf<double> d;
      BK_TRY(d.alloc(n));
      BK_CUDA(cudaMemcpyAsync(d.p, h.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
      BK_TRY(peer_allreduce_sum(p, d.p, n, st));
      BK_CUDA(cudaMemcpyAsync(h.data(), d.p, sizeof(double) * n, cudaMemcpyDeviceToHost, st));
      BK_CUDA(cudaStreamSynchronize(st));
      const double tot = 0.5 * W * (W + 1);
      for (long long i = 0; i < n; ++i)
        if (h[i] != tot * (double)((i % 97) + rep)) ++bad;
    }
  // all-gather-v of unequal segments, in place in a symmetric buffer
  {
    peer_reset(p);
    std::vector<long long> counts(W), displs(W);
    long long total = 0;
    for (int r = 0; r < W; ++r) {
      counts[r] = 1000 + 37 * r;
      displs[r] = total;
      total += counts[r];
    }
    size_t off = 0;
    BK_TRY(peer_alloc(p, sizeof(double) * total, &off));
    std::vector<double> h(total, -1.0);
    for (long long i = 0; i < counts[me]; ++i) h[displs[me] + i] = 1000.0 * me + (double)i;
    BK_TRY(peer_barrier(p, st));
    BK_CUDA(cudaMemcpyAsync(peer_ptr(p, off), h.data(), sizeof(double) * total, cudaMemcpyHostToDevice, st));
    BK_TRY(peer_barrier(p, st));  // nobody writes into a buffer its owner is still initialising
    BK_TRY(peer_allgatherv_sym(p, off, counts.data(), displs.data(), st));
    BK_CUDA(cudaMemcpyAsync(h.data(), peer_ptr(p, off), sizeof(double) * total, cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    for (int r = 0; r < W; ++r)
      for (long long i = 0; i < counts[r]; ++i)
        if (h[displs[r] + i] != 1000.0 * r + (double)i) ++bad;
    // broadcast from the last rank, a size that is not a multiple of the 4096-row blocking
    const long long nb = 3 * 4096 + 123;
    size_t ob = 0;
    BK_TRY(peer_alloc(p, sizeof(double) * nb, &ob));
    std::vector<double> hb(nb);
    for (long long i = 0; i < nb; ++i) hb[i] = (me == W - 1) ? (double)i * 0.5 : -7.0;
    BK_CUDA(cudaMemcpyAsync(peer_ptr(p, ob), hb.data(), sizeof(double) * nb, cudaMemcpyHostToDevice, st));
    BK_TRY(peer_barrier(p, st));
    BK_TRY(peer_broadcast_sym(p, ob, nb, W - 1, st));
    BK_CUDA(cudaMemcpyAsync(hb.data(), peer_ptr(p, ob), sizeof(double) * nb, cudaMemcpyDeviceToHost, st));
    BK_CUDA(cudaStreamSynchronize(st));
    for (long long i = 0; i < nb; ++i)
      if (hb[i] != (double)i * 0.5) ++bad;
    BK_TRY(peer_barrier(p, st));
  }
  BK_TRY(peer_check(p, st));
  *failures = bad;
  return BK_OK;
}

}  // extern "C"
