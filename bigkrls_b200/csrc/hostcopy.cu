// Device -> host delivery of the N x N result fields into CALLER-OWNED memory.
//
// The reference's outputs are big.matrix objects: pageable POSIX shared memory or an mmap'd backing file
// (R/bigKRLS_Rcpp_functions.R:143-147 `as.big.matrix(backingfile=)`, R/bigKRLS.R:434-453).  A plain
// cudaMemcpy into pageable memory is staged by the driver through one small pinned buffer on the calling
// thread: ~6-10 GB/s and it blocks the caller, so the "copy K under the eigensolver" overlap silently turns
// into a serial 0.4 s.  This file gives every context a small copy engine:
//
//   * pinned / registered destinations: one cudaMemcpyAsync on the copy stream (DMA at PCIe speed);
//   * pageable destinations: LANES worker threads, each an independent double-buffered pipeline over its own
//     pinned bounce buffers and its own stream - the DMA of chunk i+1 runs while the lane memcpy's chunk i
//     into the destination, and the lanes' page-faults / copies run in parallel on the host cores.
//
// Jobs are asynchronous to the submitting thread (which goes on launching the eigensolver) and ordered after
// a CUDA event on the compute stream.
#include <atomic>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#include "common.cuh"

namespace bk {

struct CopyJob {
  const char* dev = nullptr;
  char* host = nullptr;
  size_t bytes = 0;
  cudaEvent_t ready = nullptr;      // recorded on the compute stream by the submitter (may be null)
  std::atomic<size_t> next{0};      // next chunk index to claim
  std::atomic<int> active{0};       // lanes still working on this job
  std::atomic<int> error{0};
  std::mutex mu;
  std::condition_variable cv;
  bool done = false;
  // pinned destination: a single async copy, completion = this event
  cudaEvent_t pinned_done = nullptr;
};

struct HostCopier {
  static constexpr int LANES = 8;
  static constexpr size_t CHUNK = 8u << 20;  // bytes per bounce buffer
  int device = 0;
  cudaStream_t copy_stream = nullptr;
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<std::shared_ptr<CopyJob>> queue;
  bool stop = false;
  bool started = false;
  std::atomic<int> live{LANES};

  void lane_main() {
    cudaSetDevice(device);
    cudaStream_t st = nullptr;
    char* bounce[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess;
    for (int i = 0; i < 2 && ok; ++i) {
      ok = cudaHostAlloc((void**)&bounce[i], CHUNK, cudaHostAllocDefault) == cudaSuccess &&
           cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming | cudaEventBlockingSync) == cudaSuccess;
    }
    if (!ok) {
      // this lane cannot work (no pinned memory left?): leave; copier_submit falls back to a plain copy once
      // no lane is alive
      cudaGetLastError();
      live.fetch_sub(1);
      for (int i = 0; i < 2; ++i) {
        if (bounce[i]) cudaFreeHost(bounce[i]);
        if (ev[i]) cudaEventDestroy(ev[i]);
      }
      if (st) cudaStreamDestroy(st);
      return;
    }
    for (;;) {
      std::shared_ptr<CopyJob> job;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return stop || !queue.empty(); });
        if (stop && queue.empty()) break;
        job = queue.front();
        const size_t nchunks = (job->bytes + CHUNK - 1) / CHUNK;
        if (job->next.load() >= nchunks) {  // fully claimed: retire it from the queue
          queue.pop_front();
          continue;
        }
        job->active.fetch_add(1);
      }
      if (job->ready) cudaStreamWaitEvent(st, job->ready, 0);
      const size_t nchunks = (job->bytes + CHUNK - 1) / CHUNK;
      // double-buffered: issue the DMA of the next claimed chunk before draining the previous one
      size_t pend_off[2] = {0, 0}, pend_sz[2] = {0, 0};
      bool pend[2] = {false, false};
      int b = 0;
      for (;;) {
        const size_t c = job->next.fetch_add(1);
        const bool have = c < nchunks;
        if (have) {
          const size_t off = c * CHUNK, sz = std::min(CHUNK, job->bytes - off);
          if (cudaMemcpyAsync(bounce[b], job->dev + off, sz, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
              cudaEventRecord(ev[b], st) != cudaSuccess)
            job->error.store(1);
          pend[b] = true;
          pend_off[b] = off;
          pend_sz[b] = sz;
        }
        const int o = b ^ 1;
        if (pend[o]) {
          if (cudaEventSynchronize(ev[o]) != cudaSuccess) job->error.store(1);
          memcpy(job->host + pend_off[o], bounce[o], pend_sz[o]);
          pend[o] = false;
        }
        if (!have) {
          if (pend[b]) {
            if (cudaEventSynchronize(ev[b]) != cudaSuccess) job->error.store(1);
            memcpy(job->host + pend_off[b], bounce[b], pend_sz[b]);
            pend[b] = false;
          }
          break;
        }
        b = o;
      }
      if (job->active.fetch_sub(1) == 1 && job->next.load() >= nchunks) {
        std::lock_guard<std::mutex> lk(job->mu);
        job->done = true;
        job->cv.notify_all();
      }
    }
    for (int i = 0; i < 2; ++i) {
      if (bounce[i]) cudaFreeHost(bounce[i]);
      if (ev[i]) cudaEventDestroy(ev[i]);
    }
    if (st) cudaStreamDestroy(st);
  }

  void start() {
    if (started) return;
    started = true;
    for (int i = 0; i < LANES; ++i) threads.emplace_back([this] { lane_main(); });
  }

  ~HostCopier() {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    for (auto& t : threads) t.join();
  }
};

static bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

HostCopier* copier_create(int device, cudaStream_t copy_stream) {
  HostCopier* c = new HostCopier();
  c->device = device;
  c->copy_stream = copy_stream;
  return c;
}
void copier_destroy(HostCopier* c) { delete c; }

// Queues dev -> host (bytes) after `ready` (an event on the compute stream, or null = after the work queued on
// `after` so far).  Returns a ticket; copier_wait blocks until the bytes are in `host`.
int copier_submit(bk_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t after, CopyTicket* out) {
  HostCopier* c = ctx->copier;
  auto job = std::make_shared<CopyJob>();
  job->dev = (const char*)dev;
  job->host = (char*)host;
  job->bytes = bytes;
  BK_CUDA(cudaEventCreateWithFlags(&job->ready, cudaEventDisableTiming));
  BK_CUDA(cudaEventRecord(job->ready, after));
  if (is_pinned(host)) {
    BK_CUDA(cudaStreamWaitEvent(c->copy_stream, job->ready, 0));
    BK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    BK_CUDA(cudaEventCreateWithFlags(&job->pinned_done, cudaEventDisableTiming));
    BK_CUDA(cudaEventRecord(job->pinned_done, c->copy_stream));
  } else if (c->live.load() > 0) {
    std::lock_guard<std::mutex> lk(c->mu);
    c->start();
    c->queue.push_back(job);
    c->cv.notify_all();
  } else {
    // no working lane: the driver's own staged copy (blocking)
    BK_CUDA(cudaStreamWaitEvent(c->copy_stream, job->ready, 0));
    BK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->copy_stream));
    BK_CUDA(cudaStreamSynchronize(c->copy_stream));
    job->done = true;
  }
  out->job = job;
  return BK_OK;
}

int copier_wait(CopyTicket* t) {
  if (!t->job) return BK_OK;
  std::shared_ptr<CopyJob> job = std::static_pointer_cast<CopyJob>(t->job);
  t->job.reset();
  int rc = BK_OK;
  if (job->pinned_done) {
    if (cudaEventSynchronize(job->pinned_done) != cudaSuccess) rc = BK_ERR_CUDA;
    cudaEventDestroy(job->pinned_done);
  } else {
    std::unique_lock<std::mutex> lk(job->mu);
    job->cv.wait(lk, [&] { return job->done; });
    if (job->error.load()) rc = BK_ERR_CUDA;
  }
  if (job->ready) cudaEventDestroy(job->ready);
  if (rc != BK_OK) set_error("device -> host copy failed");
  return rc;
}

// synchronous convenience: dev -> host after everything queued on `after`
int copy_to_host(bk_ctx* ctx, void* host, const void* dev, size_t bytes, cudaStream_t after) {
  if (bytes < (4u << 20)) {  // small: the plain path
    BK_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, after));
    BK_CUDA(cudaStreamSynchronize(after));
    return BK_OK;
  }
  CopyTicket t;
  BK_TRY(copier_submit(ctx, host, dev, bytes, after, &t));
  return copier_wait(&t);
}

}  // namespace bk
