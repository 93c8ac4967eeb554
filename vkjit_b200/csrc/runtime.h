// runtime.h — the CUDA backend: device, stream, stream-ordered memory pool, NVRTC, kernel
// cache and launches.  Replaces the reference's Backend/Array traits and VulkanBackend
// (libs/vkjit-core/src/backend/mod.rs:8-25, backend/vulkan/mod.rs:41-208, device.rs:74-240,
// buffer.rs:24-92).  Where the reference rebuilds shader module, pipeline layout, pipeline
// cache, pipeline, command pool and fence on every eval and then waits twice
// (vulkan/mod.rs:88-208), this backend compiles a trace once (NVRTC -> sm_100a cubin), keys it
// by the structural trace hash, and a repeated trace costs one cuLaunchKernel on an
// asynchronous stream.
#pragma once
#include <atomic>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "ir.h"
#include "prims.h"
#include "program.h"

namespace vkjit {

// scatter_add privatisation: bins kept in shared memory per CTA, and the launch size from which it pays
constexpr size_t kPrivatizeMaxBytes = 192 * 1024;
constexpr size_t kPrivatizeMaxBytesNoGather = 224 * 1024;
constexpr uint64_t kPrivatizeMinLanes = 1ull << 22;
// bin-range passes: bins per pass (bytes of shared memory) with / without a gather in the trace, most passes
constexpr size_t kPassBytesGather = 128 * 1024;
constexpr size_t kPassBytesNoGather = 128 * 1024;
constexpr size_t kMaxPasses = 4;
bool scatter_passes();  // $VKJIT_SADD_PASSES (default kSaddPassesDefault)
constexpr int kSaddPassesDefault = 0;

struct HashOf {
  size_t operator()(const Hash128& h) const { return (size_t)(h.lo ^ (h.hi * 0x9E3779B97F4A7C15ull)); }
};

struct CachedKernel {
  void* module = nullptr;    // CUmodule
  void* function = nullptr;  // CUfunction
  std::vector<uint32_t> key; // verified on every hit
  uint32_t nparams = 0, nroots = 0;
  bool vectorized = true;
  uint32_t ctas_per_sm = 0;  // fused scan kernels: co-resident CTAs per SM (sizes the persistent grid)
};

struct Counters {
  std::atomic<uint64_t> cache_hits{0}, cache_misses{0}, trace_launches{0}, prim_launches{0};
  std::atomic<uint64_t> last_compile_ns{0}, last_eval_ns{0}, bytes_h2d{0}, bytes_d2h{0}, pool_bytes_live{0}, collectives{0};
  std::atomic<uint64_t> disk_hits{0};
  std::atomic<uint64_t> stream_ops{0};  // every launch / copy / collective this backend puts on its stream (never reset)
  void note_prim() { prim_launches += 1; stream_ops += 1; }
};

// NVRTC front door; usable without a device (the cubin is produced offline for sm_100a).
// Returns false and fills `log` on a compile error.
bool nvrtc_compile(const std::string& src, std::vector<char>& cubin, std::string& log);

class Backend {
 public:
  static void init(int device);     // throws VKJIT_ERR_NO_DEVICE when no usable GPU
  static void shutdown();
  static bool initialized();
  static Backend& get();            // throws VKJIT_ERR_NO_DEVICE before init
  static Counters& counters();

  int device = 0;
  int sm_count = 148;
  // The ONE stream of this backend.  Whoever wants to ENQUEUE work on it goes through enqueue_stream(), which bumps
  // Counters::stream_ops: the overlap proof of back-to-back reductions (capi.cpp: vkjit_reduce — "nothing else was put
  // on the stream since the previous reduction of the chain") then holds by construction; a new stream operation
  // cannot forget the bump (the proof used to rest on every call site remembering it).  Over-counting
  // only costs an overlap, never correctness.  wait_stream() is for synchronising / querying only.
  void* enqueue_stream() { counters().stream_ops += 1; return stream_; }
  void* wait_stream() const { return stream_; }
  prims::Scratch scratch;

  // Backend::create_array (backend/mod.rs:22): stream-ordered allocation from the pool
  Array* new_array(size_t bytes);
  void* alloc(size_t bytes);
  void free_async(void* p, size_t bytes);
  void trim();                      // return the recycled blocks to the CUDA pool
  void h2d(void* dst, const void* src, size_t bytes);
  void d2h(void* dst, const void* src, size_t bytes);  // synchronises
  void d2d(void* dst, const void* src, size_t bytes);
  void sync();
  void ensure_scan_scratch(size_t n, size_t tile = prims::kScanMinTile);

  // kernel cache keyed by trace hash
  CachedKernel* lookup(const Program& p);
  CachedKernel* compile(const Ir& ir, const Program& p);
  void launch(CachedKernel* k, uint32_t grid, uint32_t block, void** args, uint32_t smem_bytes = 0);
  void clear_cache();

  // Reductions launched back to back overlap through programmatic dependent launch (prims.cu: reduce_kernel).  A
  // reduction may skip the wait in front of its streaming phase only if nothing else was put on the stream by this
  // backend since the previous overlapping reduction (`sig` = Counters::stream_ops right after it) and its input is
  // none of the results the chain still has in flight (`outs`).
  struct ReduceChain {
    std::mutex mu;
    uint64_t sig = ~0ull;
    std::vector<const void*> outs;
  } reduce_chain;

  void* stream_ = nullptr;          // cudaStream_t (use the accessors above)

 private:
  std::mutex cache_mu_;
  std::unordered_map<Hash128, CachedKernel*, HashOf> cache_;
  void* pool_ = nullptr;            // cudaMemPool_t
  std::mutex recycle_mu_;
  std::unordered_map<size_t, std::vector<void*>> recycle_;  // exact size -> released blocks
  size_t recycle_bytes_ = 0;
};

// Host-mapped word a device-side watchdog writes when a peer exchange gave up (dist.cpp); every synchronising call of
// the backend checks it and fails with VKJIT_ERR_DIST.  nullptr: nothing to check.
void set_fault_word(volatile uint32_t* host_word);

// staging.cpp: copies between PAGEABLE host memory and the device through a pinned ring with a threaded host memcpy.
bool staged_h2d(void* dst, const void* src, size_t bytes, void* stream);  // true: sync the stream before reusing src
void staged_d2h(void* dst, const void* src, size_t bytes, void* stream);  // synchronous
void staging_shutdown();

// Evaluate the Ir's schedule (+ ids): the body of Ir::eval (internal.rs:482-525).
void eval(Ir& ir, const VarId* ids, size_t n);

// Fused trace -> reduce kernel for an unevaluated var (see runtime.cpp).
Array* eval_reduce(Ir& ir, VarId id, int red);
Array* eval_temp(Ir& ir, VarId id);
bool eval_scan(Ir& ir, int mode, const std::vector<VarId>& roots, const uint32_t* initial, uint32_t* count_dev, Array** out,
               uint64_t* n_out, const uint32_t* index_base = nullptr);

}  // namespace vkjit
