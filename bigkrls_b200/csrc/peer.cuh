// Peer-memory communicator: one process per GPU, every rank's "symmetric heap" mapped into every other
// rank's address space through CUDA IPC, so that kernels store straight into the other GPUs' HBM over
// NVLink 5 / NVSwitch and synchronise with release/acquire flags at system scope.  All collectives are
// kernels on the library stream - nothing synchronises with the host, nothing goes through Python.
//
// Used by the multi-GPU fit (fit.cu), the distributed dense->band reduction (sy2sb_dist.cu) and the distributed
// K X of the block-Krylov eigensolver (eigen_topk.cu).  torch.distributed (NCCL) is only the bootstrap: it
// carries the 64-byte IPC handles between the ranks once (bk_peer_create's `exchange` callback).
#pragma once
#include "common.cuh"

#define BK_MAX_PEERS 8

namespace bk {

struct PeerDev {  // passed by value to kernels
  int rank, world;
  char* heap[BK_MAX_PEERS];  // base of every rank's heap as mapped in THIS process (heap[rank] = own memory)
};

// control area at the start of every heap
static constexpr size_t PEER_FLAG_BYTES = 128;           // one flag per 128-byte line
static constexpr int PEER_CHANNELS = 32;
static constexpr size_t PEER_ERR_OFF = 64 * 1024;        // sticky error word
static constexpr size_t PEER_CNT_OFF = 64 * 1024 + 256;  // 64 local arrival counters (last-CTA-done)
static constexpr size_t PEER_AR_OFF = 1 << 20;           // all-reduce staging: [parity 2][src 8][AR_MAX] doubles
static constexpr int PEER_AR_MAX = 32768;
static constexpr size_t PEER_CTRL_BYTES = 8u << 20;

// channels (each has its own monotone sequence counter)
enum PeerChannel { CH_COLL = 0, CH_BARRIER = 1, CH_PANEL = 2, CH_ZGATHER = 3, CH_KRYLOV = 4, CH_GATHER = 5, CH_BCAST = 6, CH_ARLARGE = 7 };

}  // namespace bk

struct bk_peer {
  bk_ctx* ctx = nullptr;
  int rank = 0, world = 1;
  bk_exchange_fn exchange = nullptr;
  void* user = nullptr;
  bk::PeerDev dev;
  size_t heap_bytes = 0;
  size_t bump = 0;                       // next free offset (>= PEER_CTRL_BYTES)
  unsigned seq[bk::PEER_CHANNELS] = {};  // last sequence number used per channel (same on every rank)
  unsigned ar_count = 0;                 // all-reduce staging parity
  unsigned arl_count = 0;                // the same for the large all-reduce
};

namespace bk {

// ---- host API (all stream-ordered on ctx->stream unless stated) --------------------------------------------
// collective: make sure every rank's heap has at least `bytes` (grows by re-creating + re-exchanging handles;
// synchronises the device).  Resets the bump allocator.
int peer_ensure_heap(bk_peer* p, size_t bytes);
// bump allocation inside the heap; every rank must perform the same sequence of calls.  Returns the byte offset.
int peer_alloc(bk_peer* p, size_t bytes, size_t* offset);
inline void peer_reset(bk_peer* p) { p->bump = PEER_CTRL_BYTES; }
inline double* peer_ptr(bk_peer* p, size_t off) { return reinterpret_cast<double*>(p->dev.heap[p->rank] + off); }

int peer_barrier(bk_peer* p, cudaStream_t st);
// in-place sum over ranks of n doubles at `buf` (any device pointer); identical bits on every rank
int peer_allreduce_sum(bk_peer* p, double* buf, long long n, cudaStream_t st);
// the same for long vectors (multi-CTA): every rank stores its vector into slot [parity][rank] of the staging area at
// heap offset stage_off on every rank (2 x world x slot_elems doubles, slot_elems >= n), waits for all, adds the
// slots in rank order
int peer_allreduce_sum_large(bk_peer* p, double* buf, long long n, size_t stage_off, long long slot_elems, cudaStream_t st);
// every rank owns the segment [displs[r], displs[r]+counts[r]) (doubles) of the symmetric buffer at heap offset
// `off`; afterwards every rank holds all segments
int peer_allgatherv_sym(bk_peer* p, size_t off, const long long* counts, const long long* displs, cudaStream_t st);
// n doubles at heap offset `off` from `root` to everyone
int peer_broadcast_sym(bk_peer* p, size_t off, long long n, int root, cudaStream_t st);
// 2-D block copy into the heaps of the ranks in dst_mask (bit r; may include self) followed by the flag `seq` on
// channel `ch` (seq = peer_next_seq on EVERY rank, pushing or not, so that the channel counters stay equal).
// src is any local device pointer.
int peer_push2d(bk_peer* p, const double* src, long long lds, int rows, int cols, size_t dst_off, long long ldd,
                unsigned dst_mask, int ch, unsigned seq, cudaStream_t st);
// wait until the ranks in src_mask have signalled `seq` on channel `ch`
int peer_wait(bk_peer* p, int ch, unsigned src_mask, unsigned seq, cudaStream_t st);
// next sequence number of a channel (for kernels that signal themselves)
inline unsigned peer_next_seq(bk_peer* p, int ch) { return ++p->seq[ch]; }
// sticky device-side error (a wait timed out): returns BK_ERR_COMM if set.  Synchronises the stream.
int peer_check(bk_peer* p, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned* peer_flag(const PeerDev& pd, int dst, int ch, int src) {
  return reinterpret_cast<unsigned*>(pd.heap[dst] + (size_t)(ch * BK_MAX_PEERS + src) * PEER_FLAG_BYTES);
}
__device__ __forceinline__ unsigned* peer_err(const PeerDev& pd) {
  return reinterpret_cast<unsigned*>(pd.heap[pd.rank] + PEER_ERR_OFF);
}
__device__ __forceinline__ unsigned* peer_counter(const PeerDev& pd, int i) {
  return reinterpret_cast<unsigned*>(pd.heap[pd.rank] + PEER_CNT_OFF) + i;
}
// one thread: spin until rank `src` has signalled `seq` on `ch`.  Gives up after ~10 s (or at once when another
// wait already failed) and leaves a sticky error word instead of hanging the GPU.
__device__ __forceinline__ void peer_wait_flag(const PeerDev& pd, int ch, int src, unsigned seq) {
  const unsigned* f = peer_flag(pd, pd.rank, ch, src);
  unsigned* err = peer_err(pd);
  const long long t0 = clock64();
  while ((int)(ld_acquire_sys_u32(f) - seq) < 0) {
    if (*reinterpret_cast<volatile unsigned*>(err) != 0u) break;
    if (clock64() - t0 > 20000000000LL) {
      atomicExch(err, 0x100u + (unsigned)src);
      break;
    }
  }
}
// Call from ALL threads of every CTA of a kernel after its last peer store: the CTA that arrives last raises the
// flag `seq` on channel `ch` at every rank of dst_mask.  `cnt` is a local arrival counter (zero between kernels).
__device__ __forceinline__ void peer_signal_last_cta(const PeerDev& pd, unsigned* cnt, unsigned total_ctas,
                                                     unsigned dst_mask, int ch, unsigned seq) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(cnt, 1u);
    if (prev == total_ctas - 1) {
      __threadfence_system();
      *cnt = 0u;
      for (int r = 0; r < pd.world; ++r)
        if (dst_mask & (1u << r)) st_release_sys_u32(peer_flag(pd, r, ch, pd.rank), seq);
    }
  }
}
#endif

}  // namespace bk
