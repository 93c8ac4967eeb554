// staging.cpp — host <-> device copies for PAGEABLE caller memory through a pinned staging ring.
//
// The reference's arrays are host-visible mapped Vulkan buffers: upload is a memcpy into mapped memory
// (backend/vulkan/mod.rs:56-73) and readback a zero-copy view (Ir::as_slice, internal.rs:443-449; Array::map,
// vulkan/mod.rs:29-34).  HBM is not host-visible, so every upload / readback crosses PCIe.  A cudaMemcpy straight
// from / to pageable memory makes the driver stage through its own small pinned buffers with ONE thread
// (measured round 1: 4 MiB readback 252 us pageable vs 98 us pinned; 1 GiB upload ~3x slower than pinned).  This ring
// does the staging itself: kChunks pinned chunks (8 MiB each), the DMA of chunk i overlapping the host memcpy of chunk i +- 1, and the
// host memcpy split over a few worker threads (one core copies ~10 GB/s, PCIe 5 x16 moves ~55).
//
// Transfers below kRingMinBytes keep the driver's own pageable path: waking the copy threads costs ~0.1 ms, more than the
// ring saves there (measured: 4 MiB readback 252 us through the driver, 720 us through the ring with cold workers;
// 1 GiB upload 32 GB/s through the ring vs ~18 GB/s through the driver).  The front-ends avoid the question for reads:
// their result arrays live in pinned memory (vkjit_b200/ir.py: as_slice), one DMA, 98-110 us for 4 MiB.
//
// H2D returns as soon as the caller's buffer has been copied OUT (the data waits in pinned chunks; a chunk is reused only
// after the event behind its DMA) — the reference's contract "the slice may be reused when the call returns" without a
// device synchronisation.  Pinned caller memory (vkjit_host_alloc, cudaHostRegister) skips the ring: one direct DMA.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "runtime.h"

namespace vkjit {
namespace {

void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

// memcpy split over worker threads.  The workers spin briefly before they sleep: copies come in bursts (one per chunk).
class CopyPool {
 public:
  explicit CopyPool(int workers) {
    for (int i = 0; i < workers; ++i) threads_.emplace_back([this, i] { run(i); });
  }
  ~CopyPool() {
    { std::lock_guard<std::mutex> g(mu_); stop_ = true; ++generation_; gen_atomic_.store(generation_, std::memory_order_release); }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void copy(void* dst, const void* src, size_t bytes) {
    const size_t parts = threads_.size() + 1;
    if (bytes < (256u << 10) || parts == 1) { memcpy(dst, src, bytes); return; }
    const size_t per = ((bytes / parts) + 4095) & ~(size_t)4095;
    {
      std::lock_guard<std::mutex> g(mu_);
      dst_ = (char*)dst; src_ = (const char*)src; bytes_ = bytes; per_ = per;
      pending_.store((int)threads_.size(), std::memory_order_relaxed);
      ++generation_;
      gen_atomic_.store(generation_, std::memory_order_release);
    }
    cv_.notify_all();
    part(0);
    while (pending_.load(std::memory_order_acquire) != 0) std::this_thread::yield();
  }

 private:
  void part(size_t k) {
    const size_t lo = std::min(bytes_, k * per_), hi = (k + 1 == threads_.size() + 1) ? bytes_ : std::min(bytes_, lo + per_);
    if (hi > lo) memcpy(dst_ + lo, src_ + lo, hi - lo);
  }
  void run(int i) {
    uint64_t seen = 0;
    for (;;) {
      // spin briefly before sleeping: the chunks of one transfer follow each other within tens of microseconds, a
      // condition-variable wake-up alone costs about as much as copying a chunk
      bool got = false;
      for (int spin = 0; spin < 20000 && !got; ++spin) {
        if (gen_atomic_.load(std::memory_order_acquire) != seen) got = true;
        else if ((spin & 63) == 63) std::this_thread::yield();
      }
      {
        std::unique_lock<std::mutex> lk(mu_);
        if (!got) cv_.wait(lk, [&] { return generation_ != seen; });
        seen = generation_;
        if (stop_) return;
      }
      part((size_t)i + 1);
      pending_.fetch_sub(1, std::memory_order_release);
    }
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  uint64_t generation_ = 0;
  std::atomic<uint64_t> gen_atomic_{0};
  bool stop_ = false;
  char* dst_ = nullptr; const char* src_ = nullptr; size_t bytes_ = 0, per_ = 0;
  std::atomic<int> pending_{0};
};

constexpr int kChunks = 4;
constexpr size_t kRingMinBytes = 16u << 20;

struct Ring {
  size_t chunk_bytes = 0;
  void* chunk[kChunks] = {};
  cudaEvent_t ev[kChunks] = {};
  bool busy[kChunks] = {};
  CopyPool* pool = nullptr;
  std::mutex mu;  // one transfer at a time (the Ir lock does not cover two Irs on one backend)
};
Ring* g_ring = nullptr;
std::mutex g_ring_mu;

Ring& ring() {
  std::lock_guard<std::mutex> g(g_ring_mu);
  if (!g_ring) {
    auto* r = new Ring();
    const char* cb = getenv("VKJIT_STAGING_CHUNK_KB");
    r->chunk_bytes = (size_t)(cb ? std::max(64, atoi(cb)) : 8192) << 10;  // 8 MiB: measured optimum (profiles/r02_staging.md: 2 MiB 33.5, 8 MiB 50.7, 16 MiB 36.6 GB/s H2D from pageable memory)
    for (int i = 0; i < kChunks; ++i) {
      ck(cudaHostAlloc(&r->chunk[i], r->chunk_bytes, cudaHostAllocDefault), "staging chunk allocation");
      ck(cudaEventCreateWithFlags(&r->ev[i], cudaEventDisableTiming), "staging event");
    }
    const char* ct = getenv("VKJIT_COPY_THREADS");
    int threads = ct ? atoi(ct) : (int)std::min(8u, std::max(1u, std::thread::hardware_concurrency() / 2));
    r->pool = new CopyPool(std::max(0, threads - 1));
    g_ring = r;
  }
  return *g_ring;
}

bool is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost;
}

}  // namespace

void staging_shutdown() {
  std::lock_guard<std::mutex> g(g_ring_mu);
  if (!g_ring) return;
  delete g_ring->pool;
  for (int i = 0; i < kChunks; ++i) { cudaFreeHost(g_ring->chunk[i]); cudaEventDestroy(g_ring->ev[i]); }
  delete g_ring;
  g_ring = nullptr;
}

// returns true if the stream still has to be synchronised before `src` may be reused (direct DMA from pinned memory)
bool staged_h2d(void* dst, const void* src, size_t bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  static const bool off = getenv("VKJIT_NO_STAGING") != nullptr;
  if (off || bytes < kRingMinBytes || is_pinned(src)) {
    ck(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s), "H2D copy");
    return true;
  }
  Ring& r = ring();
  std::lock_guard<std::mutex> g(r.mu);
  size_t off_b = 0;
  for (int i = 0; off_b < bytes; ++i) {
    const int slot = i % kChunks;
    const size_t len = std::min(r.chunk_bytes, bytes - off_b);
    if (r.busy[slot]) ck(cudaEventSynchronize(r.ev[slot]), "staging wait");  // its previous DMA has read the chunk
    r.pool->copy(r.chunk[slot], (const char*)src + off_b, len);
    ck(cudaMemcpyAsync((char*)dst + off_b, r.chunk[slot], len, cudaMemcpyHostToDevice, s), "H2D copy");
    ck(cudaEventRecord(r.ev[slot], s), "staging event");
    r.busy[slot] = true;
    off_b += len;
  }
  return false;  // the caller's buffer has been copied out completely
}

// synchronous: dst holds the data on return
void staged_d2h(void* dst, const void* src, size_t bytes, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
  static const bool off = getenv("VKJIT_NO_STAGING") != nullptr;
  if (off || bytes < kRingMinBytes || is_pinned(dst)) {
    if (bytes) ck(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s), "D2H copy");
    ck(cudaStreamSynchronize(s), "D2H sync");
    return;
  }
  Ring& r = ring();
  std::lock_guard<std::mutex> g(r.mu);
  for (int k = 0; k < kChunks; ++k)
    if (r.busy[k]) { ck(cudaEventSynchronize(r.ev[k]), "staging wait"); r.busy[k] = false; }  // earlier uploads are done with the chunks
  const size_t nchunks = (bytes + r.chunk_bytes - 1) / r.chunk_bytes;
  auto issue = [&](size_t i) {
    const size_t o = i * r.chunk_bytes, len = std::min(r.chunk_bytes, bytes - o);
    ck(cudaMemcpyAsync(r.chunk[i % kChunks], (const char*)src + o, len, cudaMemcpyDeviceToHost, s), "D2H copy");
    ck(cudaEventRecord(r.ev[i % kChunks], s), "staging event");
  };
  for (size_t i = 0; i < nchunks && i < (size_t)kChunks; ++i) issue(i);
  for (size_t i = 0; i < nchunks; ++i) {
    const size_t o = i * r.chunk_bytes, len = std::min(r.chunk_bytes, bytes - o);
    ck(cudaEventSynchronize(r.ev[i % kChunks]), "staging wait");
    r.pool->copy((char*)dst + o, r.chunk[i % kChunks], len);   // the DMA of chunks i+1 .. i+3 runs meanwhile
    if (i + kChunks < nchunks) issue(i + kChunks);
  }
}

}  // namespace vkjit
