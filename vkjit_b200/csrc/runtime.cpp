// runtime.cpp — CUDA backend (see runtime.h).
#include "runtime.h"

#include <cuda.h>
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>

namespace vkjit {

uint64_t now_ns() {
  return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

namespace {

// libcuda is bound at run time so that the library loads (and traces can be built, hashed and
// compiled) on a machine without a driver; cudart is linked statically.
struct Driver {
  void* handle = nullptr;
  CUresult (*ModuleLoadData)(CUmodule*, const void*) = nullptr;
  CUresult (*ModuleUnload)(CUmodule) = nullptr;
  CUresult (*ModuleGetFunction)(CUfunction*, CUmodule, const char*) = nullptr;
  CUresult (*LaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**) = nullptr;
  CUresult (*GetErrorString)(CUresult, const char**) = nullptr;
  CUresult (*FuncSetAttribute)(CUfunction, CUfunction_attribute, int) = nullptr;
  CUresult (*OccupancyMaxActiveBlocks)(int*, CUfunction, int, size_t) = nullptr;
  bool load(std::string& why) {
    if (handle) return true;
    handle = dlopen("libcuda.so.1", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) { why = std::string("cannot load libcuda.so.1: ") + dlerror(); return false; }
    auto sym = [&](const char* n) { return dlsym(handle, n); };
    ModuleLoadData = (decltype(ModuleLoadData))sym("cuModuleLoadData");
    ModuleUnload = (decltype(ModuleUnload))sym("cuModuleUnload");
    ModuleGetFunction = (decltype(ModuleGetFunction))sym("cuModuleGetFunction");
    LaunchKernel = (decltype(LaunchKernel))sym("cuLaunchKernel");
    GetErrorString = (decltype(GetErrorString))sym("cuGetErrorString");
    FuncSetAttribute = (decltype(FuncSetAttribute))sym("cuFuncSetAttribute");
    OccupancyMaxActiveBlocks = (decltype(OccupancyMaxActiveBlocks))sym("cuOccupancyMaxActiveBlocksPerMultiprocessor");
    if (!OccupancyMaxActiveBlocks || !ModuleLoadData || !ModuleUnload || !ModuleGetFunction || !LaunchKernel || !GetErrorString || !FuncSetAttribute) {
      why = "libcuda.so.1 lacks required entry points";
      return false;
    }
    return true;
  }
};

Driver g_drv;
Backend* g_backend = nullptr;
std::mutex g_init_mu;
Counters g_counters;
std::atomic<volatile uint32_t*> g_fault_word{nullptr};

// after a stream synchronisation: did an exchange watchdog fire while the stream ran?
void check_fault() {
  volatile uint32_t* w = g_fault_word.load(std::memory_order_relaxed);
  if (w && *w) {
    const uint32_t seq = *w;
    *w = 0;
    fail(VKJIT_ERR_DIST, "multi-GPU exchange " + std::to_string(seq) + ": a peer did not arrive within the watchdog ($VKJIT_DIST_TIMEOUT_S); "
                         "results produced since are unspecified");
  }
}

void ck(cudaError_t e, const char* what) {
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
void cku(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS) {
    const char* s = nullptr;
    if (g_drv.GetErrorString) g_drv.GetErrorString(r, &s);
    fail(VKJIT_ERR_CUDA, std::string(what) + ": " + (s ? s : "unknown driver error"));
  }
}

}  // namespace

void set_fault_word(volatile uint32_t* host_word) { g_fault_word.store(host_word); }

Counters& Backend::counters() { return g_counters; }
bool Backend::initialized() { return g_backend != nullptr; }

Backend& Backend::get() {
  if (!g_backend)
    fail(VKJIT_ERR_NO_DEVICE, "vkjit_b200 backend is not initialised (call vkjit_init on a machine with a B200); there is no CPU fallback");
  return *g_backend;
}

void Backend::init(int device) {
  std::lock_guard<std::mutex> g(g_init_mu);
  if (g_backend) return;
  if (device < 0) {
    const char* lr = getenv("LOCAL_RANK");
    device = lr ? atoi(lr) : 0;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    fail(VKJIT_ERR_NO_DEVICE, std::string("no CUDA device available (") + (e != cudaSuccess ? cudaGetErrorString(e) : "0 devices") +
                                  "); vkjit_b200 has no CPU fallback");
  if (device >= count) fail(VKJIT_ERR_NO_DEVICE, "device index " + std::to_string(device) + " out of range");
  std::string why;
  if (!g_drv.load(why)) fail(VKJIT_ERR_NO_DEVICE, why);
  ck(cudaSetDevice(device), "cudaSetDevice");
  ck(cudaFree(nullptr), "context creation");
  int major = 0, minor = 0, sms = 0;
  ck(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device), "attr");
  ck(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device), "attr");
  ck(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), "attr");
  if (major != 10)
    fail(VKJIT_ERR_NO_DEVICE, "device compute capability " + std::to_string(major) + "." + std::to_string(minor) +
                                  " is not sm_100: this backend only generates sm_100a (B200) code");
  auto* b = new Backend();
  b->device = device;
  b->sm_count = sms;
  cudaStream_t s;
  ck(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
  b->stream_ = s;
  cudaMemPool_t pool;
  ck(cudaDeviceGetDefaultMemPool(&pool, device), "cudaDeviceGetDefaultMemPool");
  uint64_t keep = ~0ull;  // never trim: freed blocks stay in the pool for the next eval
  ck(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep), "cudaMemPoolSetAttribute");
  b->pool_ = pool;
  void* p = nullptr;
  ck(cudaMalloc(&p, (size_t)prims::kReduceMaxCtas * 4), "scratch");
  b->scratch.partials = p;
  ck(cudaMalloc(&p, 256), "scratch");
  ck(cudaMemset(p, 0, 256), "scratch");
  b->scratch.ticket = (unsigned int*)p;
  g_backend = b;
}

void Backend::shutdown() {
  std::lock_guard<std::mutex> g(g_init_mu);
  if (!g_backend) return;
  Backend* b = g_backend;
  staging_shutdown();
  b->trim();
  cudaStreamSynchronize((cudaStream_t)b->stream_);
  b->clear_cache();
  cudaFree(b->scratch.partials);
  cudaFree(b->scratch.ticket);
  if (b->scratch.tile_state) cudaFree(b->scratch.tile_state);
  cudaStreamDestroy((cudaStream_t)b->stream_);
  g_backend = nullptr;
  delete b;
}

// Small blocks are recycled by exact size in front of the CUDA pool.  Every kernel, copy and collective of this
// backend runs on the ONE backend stream, so a block released after its last user was enqueued may be handed to
// the next request at once — the same guarantee cudaFreeAsync/cudaMallocAsync give, minus two driver calls
// (~1 us of the ~3.3 us a cached trace launch costs).  Bounded: blocks <= 4 MiB, <= 8 per size, <= 64 MiB in all.
namespace {
constexpr size_t kRecycleMaxBlock = 4u << 20, kRecycleMaxPerSize = 8, kRecycleMaxTotal = 64u << 20;
}

void* Backend::alloc(size_t bytes) {
  void* p = nullptr;
  const size_t sz = bytes ? bytes : 16;
  if (sz <= kRecycleMaxBlock) {
    std::lock_guard<std::mutex> g(recycle_mu_);
    auto it = recycle_.find(sz);
    if (it != recycle_.end() && !it->second.empty()) {
      p = it->second.back();
      it->second.pop_back();
      recycle_bytes_ -= sz;
    }
  }
  if (!p) ck(cudaMallocAsync(&p, sz, (cudaStream_t)stream_), "cudaMallocAsync");
  g_counters.pool_bytes_live += sz;
  return p;
}

void Backend::free_async(void* p, size_t bytes) {
  if (!p) return;
  const size_t sz = bytes ? bytes : 16;
  g_counters.pool_bytes_live -= sz;
  if (sz <= kRecycleMaxBlock) {
    std::lock_guard<std::mutex> g(recycle_mu_);
    std::vector<void*>& slot = recycle_[sz];
    if (slot.size() < kRecycleMaxPerSize && recycle_bytes_ + sz <= kRecycleMaxTotal) {
      slot.push_back(p);
      recycle_bytes_ += sz;
      return;
    }
  }
  cudaFreeAsync(p, (cudaStream_t)stream_);
}

void Backend::trim() {
  std::lock_guard<std::mutex> g(recycle_mu_);
  for (auto& kv : recycle_)
    for (void* p : kv.second) cudaFreeAsync(p, (cudaStream_t)stream_);
  recycle_.clear();
  recycle_bytes_ = 0;
}

Array* Backend::new_array(size_t bytes) {
  Array* a = new Array();
  a->ptr = alloc(bytes);
  a->bytes = bytes;
  a->capacity = bytes;
  return a;
}

namespace {
thread_local std::vector<std::pair<void (*)(void*), void*>> g_pending_release;
}

void release_array(Array* a) {
  if (!a) return;
  if (a->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) return;  // an exported tensor (or the var) still holds it
  if (g_backend && a->owned) {
    if (a->exposed) {
      // Whoever received the pointer (torch, cupy, a raw device_ptr user) may have work on ITS streams that still
      // touches the memory; the stream-ordered recycler / pool only order against the backend stream (and cudaFree of
      // a stream-ordered allocation does not synchronise either).  Rare path (exports), correctness over speed.
      cudaDeviceSynchronize();
      g_backend->free_async(a->ptr, a->capacity);
    } else {
      g_backend->free_async(a->ptr, a->capacity);
    }
  }
  if (a->release) g_pending_release.emplace_back(a->release, a->release_ctx);
  delete a;
}

void drain_foreign_releases() {
  while (!g_pending_release.empty()) {
    auto pr = g_pending_release.back();
    g_pending_release.pop_back();
    // the consumer may still have work queued on our stream that reads the memory
    if (g_backend) cudaStreamSynchronize((cudaStream_t)g_backend->stream_);
    pr.first(pr.second);
  }
}

void Backend::h2d(void* dst, const void* src, size_t bytes) {
  if (!bytes) return;
  // the caller may reuse `src` as soon as we return (Backend::create_array_from_slice copies synchronously into
  // mapped memory, vulkan/mod.rs:56-73): pageable memory is copied out into the pinned ring (no device
  // synchronisation), a direct DMA from pinned memory has to finish first
  if (staged_h2d(dst, src, bytes, stream_)) ck(cudaStreamSynchronize((cudaStream_t)stream_), "H2D sync");
  g_counters.bytes_h2d += bytes;
  g_counters.stream_ops += 1;
}

void Backend::d2h(void* dst, const void* src, size_t bytes) {
  staged_d2h(dst, src, bytes, stream_);
  check_fault();
  g_counters.bytes_d2h += bytes;
  g_counters.stream_ops += 1;
}

void Backend::d2d(void* dst, const void* src, size_t bytes) {
  if (bytes) ck(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream_), "D2D copy");
  g_counters.stream_ops += 1;
}

void Backend::sync() {
  ck(cudaStreamSynchronize((cudaStream_t)stream_), "cudaStreamSynchronize");
  check_fault();
}

void Backend::ensure_scan_scratch(size_t n, size_t tile) {
  const size_t need = prims::scan_state_words(n, tile);
  if (need <= scratch.tile_state_words) return;
  if (scratch.tile_state) {
    ck(cudaStreamSynchronize((cudaStream_t)stream_), "sync");
    cudaFree(scratch.tile_state);
  }
  size_t words = std::max<size_t>(need, 1 << 19);
  void* p = nullptr;
  ck(cudaMalloc(&p, words * 8), "scan scratch");
  scratch.tile_state = (uint64_t*)p;
  scratch.tile_state_words = words;
}

// ---- NVRTC -----------------------------------------------------------------------------
namespace {
// (also hashed into the disk cache header: a different option set is a different cubin)
const char* const kNvrtcOptions[] = {"--gpu-architecture=sm_100a", "--fmad=false", "--prec-div=true", "--prec-sqrt=true", "--ftz=false",
                                     "-lineinfo",                  "--std=c++17",  "-default-device", "--minimal"};
}  // namespace

bool nvrtc_compile(const std::string& src, std::vector<char>& cubin, std::string& log) {
  nvrtcProgram prog;
  if (nvrtcCreateProgram(&prog, src.c_str(), "vkjit_trace.cu", 0, nullptr, nullptr) != NVRTC_SUCCESS) {
    log = "nvrtcCreateProgram failed";
    return false;
  }
  // --fmad=false: the reference's OpFMul/OpFAdd are separate roundings (no contraction); f32
  // + - * / are additionally emitted as __f*_rn intrinsics, which never contract.  IEEE
  // division/sqrt and no flush-to-zero are the NVRTC defaults and are stated explicitly.
  // --minimal: leaves texture / surface / cudadevrt declarations out of the implicit header — 20 % less compile time
  // (3-op trace 84 -> 66 ms, 364-node trace 256 -> 200 ms on the build container's CPU) for byte-identical cubins
  // (checked for every kernel family the generator emits).
  nvrtcResult r = nvrtcCompileProgram(prog, (int)(sizeof(kNvrtcOptions) / sizeof(kNvrtcOptions[0])), kNvrtcOptions);
  size_t ls = 0;
  nvrtcGetProgramLogSize(prog, &ls);
  if (ls > 1) { log.resize(ls); nvrtcGetProgramLog(prog, &log[0]); }
  if (r != NVRTC_SUCCESS) {
    if (log.empty()) log = nvrtcGetErrorString(r);
    nvrtcDestroyProgram(&prog);
    return false;
  }
  size_t cs = 0;
  if (nvrtcGetCUBINSize(prog, &cs) != NVRTC_SUCCESS || cs == 0) {
    log += "\nnvrtcGetCUBINSize failed";
    nvrtcDestroyProgram(&prog);
    return false;
  }
  cubin.resize(cs);
  nvrtcGetCUBIN(prog, cubin.data());
  nvrtcDestroyProgram(&prog);
  if (const char* dir = getenv("VKJIT_DUMP_CUBIN_DIR")) {  // for cuobjdump -sass (profiles/r02_sass.md); needs no device
    static std::atomic<int> seq{0};
    const std::string base = std::string(dir) + "/vkjit_" + std::to_string(seq++);
    if (FILE* f = fopen((base + ".cubin").c_str(), "wb")) { fwrite(cubin.data(), 1, cubin.size(), f); fclose(f); }
    if (FILE* f = fopen((base + ".cu").c_str(), "wb")) { fwrite(src.data(), 1, src.size(), f); fclose(f); }
  }
  return true;
}

// ---- on-disk cubin cache --------------------------------------------------
// $VKJIT_CACHE_DIR/<hash>.cubin = {magic, nvrtc version, generator fingerprint ^ options hash, key_len, key words,
// cubin}.  The canonical key is stored and compared, so a hash collision can never load the wrong kernel, and the
// fingerprint (program.cpp: embedded device sources + generator revision, plus the NVRTC option string) makes a
// cubin written by a different build of this library a plain miss.
namespace {
constexpr uint32_t kDiskMagic = 0x564B4333u;  // "VKC3" (header carries the generator fingerprint)
uint32_t build_fingerprint() {
  uint32_t h = generator_fingerprint();
  for (const char* o : kNvrtcOptions)
    for (const char* t = o; *t; ++t) { h ^= (unsigned char)*t; h *= 16777619u; }
  return h;
}

std::string disk_path(const Hash128& h) {
  const char* dir = getenv("VKJIT_CACHE_DIR");
  if (!dir || !*dir) return std::string();
  char name[64];
  snprintf(name, sizeof name, "/%016llx%016llx.cubin", (unsigned long long)h.hi, (unsigned long long)h.lo);
  return std::string(dir) + name;
}

uint32_t nvrtc_version_word() {
  int major = 0, minor = 0;
  nvrtcVersion(&major, &minor);
  return (uint32_t)(major * 1000 + minor);
}

bool disk_load(const Program& p, std::vector<char>& cubin) {
  const std::string path = disk_path(p.hash);
  if (path.empty()) return false;
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  bool ok = false;
  uint32_t hdr[4];
  if (fread(hdr, 4, 4, f) == 4 && hdr[0] == kDiskMagic && hdr[1] == nvrtc_version_word() && hdr[2] == build_fingerprint() &&
      hdr[3] == p.key_len) {
    std::vector<uint32_t> key(hdr[3]);
    if (fread(key.data(), 4, key.size(), f) == key.size() && memcmp(key.data(), p.key.data(), key.size() * 4) == 0) {
      const long at = ftell(f);
      fseek(f, 0, SEEK_END);
      const long end = ftell(f);
      fseek(f, at, SEEK_SET);
      if (end > at) {
        cubin.resize((size_t)(end - at));
        ok = fread(cubin.data(), 1, cubin.size(), f) == cubin.size();
      }
    }
  }
  fclose(f);
  return ok;
}

void disk_store(const Program& p, const std::vector<char>& cubin) {
  const std::string path = disk_path(p.hash);
  if (path.empty()) return;
  const std::string tmp = path + ".tmp" + std::to_string((unsigned long long)now_ns());
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  const uint32_t hdr[4] = {kDiskMagic, nvrtc_version_word(), build_fingerprint(), (uint32_t)p.key_len};
  bool ok = fwrite(hdr, 4, 4, f) == 4 && fwrite(p.key.data(), 4, p.key_len, f) == p.key_len &&
            fwrite(cubin.data(), 1, cubin.size(), f) == cubin.size();
  fclose(f);
  if (ok) rename(tmp.c_str(), path.c_str());  // atomic publish
  else remove(tmp.c_str());
}
}  // namespace

// ---- kernel cache ------------------------------------------------------------------------
CachedKernel* Backend::lookup(const Program& p) {
  std::lock_guard<std::mutex> g(cache_mu_);
  auto it = cache_.find(p.hash);
  if (it == cache_.end()) return nullptr;
  CachedKernel* k = it->second;
  if (k->key.size() != p.key_len || memcmp(k->key.data(), p.key.data(), p.key_len * 4) != 0) return nullptr;  // 128-bit collision
  g_counters.cache_hits += 1;
  return k;
}

CachedKernel* Backend::compile(const Ir& ir, const Program& p) {
  const uint64_t t0 = now_ns();
  std::vector<char> cubin;
  if (disk_load(p, cubin)) {
    g_counters.disk_hits += 1;
  } else {
    const std::string src = generate_cuda(ir, p);
    if (getenv("VKJIT_DUMP")) fprintf(stderr, "%s\n", src.c_str());
    std::string log;
    if (!nvrtc_compile(src, cubin, log)) fail(VKJIT_ERR_COMPILE, "NVRTC rejected the generated kernel:\n" + log + "\n--- source ---\n" + src);
    disk_store(p, cubin);
  }
  CUmodule mod;
  cku(g_drv.ModuleLoadData(&mod, cubin.data()), "cuModuleLoadData");
  CUfunction fn;
  cku(g_drv.ModuleGetFunction(&fn, mod, "vkjit_trace"), "cuModuleGetFunction");
  if (p.privatize)
    cku(g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)kPrivatizeMaxBytesNoGather), "cuFuncSetAttribute");
  int ctas_per_sm = 0;
  if (p.scan >= 0) {
    const ScanFusedGeom sg = scan_fused_geom(stream_count(p), p.scan, p.order.size());
    const size_t smem = sg.smem(stream_count(p));
    if (smem) cku(g_drv.FuncSetAttribute(fn, CU_FUNC_ATTRIBUTE_MAX_DYNAMIC_SHARED_SIZE_BYTES, (int)smem), "cuFuncSetAttribute");
    // the look-back needs every CTA of the grid co-resident: check the real occupancy
    cku(g_drv.OccupancyMaxActiveBlocks(&ctas_per_sm, fn, sg.launch_threads(), smem), "cuOccupancyMaxActiveBlocksPerMultiprocessor");
    if (ctas_per_sm < 1) fail(VKJIT_ERR_CUDA, "fused scan kernel does not fit on an SM");
    if (sg.ctrl) ctas_per_sm = std::min(ctas_per_sm, sg.ctas);
    else ctas_per_sm = std::min(ctas_per_sm, 1024 / sg.threads);  // one 1024-thread CTA or two 512-thread CTAs per SM (64 registers per thread)
  }
  auto* k = new CachedKernel();
  k->ctas_per_sm = (uint32_t)ctas_per_sm;
  k->module = mod; k->function = fn; k->key.assign(p.key.begin(), p.key.begin() + p.key_len);
  k->nparams = (uint32_t)p.params.size(); k->nroots = (uint32_t)p.roots.size(); k->vectorized = p.vectorized;
  {
    std::lock_guard<std::mutex> g(cache_mu_);
    auto it = cache_.find(p.hash);
    if (it != cache_.end()) {  // hash collision with a different key: replace
      g_drv.ModuleUnload((CUmodule)it->second->module);
      delete it->second;
      cache_.erase(it);
    }
    cache_[p.hash] = k;
  }
  g_counters.cache_misses += 1;
  g_counters.last_compile_ns = now_ns() - t0;
  return k;
}

void Backend::clear_cache() {
  std::lock_guard<std::mutex> g(cache_mu_);
  cudaStreamSynchronize((cudaStream_t)stream_);
  for (auto& kv : cache_) {
    if (g_drv.ModuleUnload) g_drv.ModuleUnload((CUmodule)kv.second->module);
    delete kv.second;
  }
  cache_.clear();
}

void Backend::launch(CachedKernel* k, uint32_t grid, uint32_t block, void** args, uint32_t smem_bytes) {
  cku(g_drv.LaunchKernel((CUfunction)k->function, grid, 1, 1, block, 1, 1, smem_bytes, (CUstream)stream_, args, nullptr), "cuLaunchKernel");
  g_counters.trace_launches += 1;
  g_counters.stream_ops += 1;
}

// ---- Ir::eval (internal.rs:482-525) ---------------------------------------------------------
bool scatter_passes() {
  static const int on = [] { const char* s = getenv("VKJIT_SADD_PASSES"); return s ? (s[0] == '1' ? 1 : 0) : kSaddPassesDefault; }();
  return on == 1;
}
namespace {

// Compile-or-lookup + launch of one group of roots that share a kernel size.  commit: the roots become
// Bindings (Ir::eval); otherwise the fresh arrays are handed to the caller and the vars stay as they are.
// $VKJIT_EVAL_TRACE=1: per-phase host time of every eval on stderr (walk / lookup / alloc / launch / commit)
bool eval_trace_on() {
  static const bool on = getenv("VKJIT_EVAL_TRACE") != nullptr;
  return on;
}

void run_group(Ir& ir, Backend& be, const std::vector<VarId>& roots, bool commit, std::vector<Array*>* result) {
  const bool trace = eval_trace_on();
  uint64_t ts[6] = {0, 0, 0, 0, 0, 0};
  if (trace) ts[0] = now_ns();
  static thread_local Program prog;
  static thread_local std::vector<Array*> outs;
  static thread_local std::vector<void*> argv;
  static thread_local std::vector<uint64_t> ptrs;
  outs.clear();
  try {
    build_program(ir, roots, true, prog);
    if (prog.n == 0) fail(VKJIT_ERR_SIZE, "zero-sized kernel");
    bool aligned = true;
    for (const Param& pr : prog.params)
      if ((pr.use & USE_STREAM) && ((uintptr_t)ir.vars[pr.var].array->ptr & 15u)) aligned = false;
    // scatter_add with many lanes: keep as many bins of the target as fit in shared memory (zeroing and
    // flushing them costs ~2 x kbins atomics per CTA, so small launches keep the plain L2 path)
    uint32_t kbins = 0;
    bool cluster_bins = false;
    if (prog.sadd_param >= 0 && prog.n >= kPrivatizeMinLanes && !getenv("VKJIT_NO_PRIVATIZE")) {
      const size_t bins = ir.vars[prog.params[prog.sadd_param].var].array->bytes / 4;
      // measured on H26 (profiles/hist_ab.py): with a gather in the trace 192 KB of bins is the optimum (more
      // shared memory leaves too little L1 for the gather table: 0.41 ms at 192 KB, 0.55 ms at 224 KB); without
      // one, more is better (counting histogram: 0.25 ms at 192 KB, 0.19 ms at 224 KB)
      size_t max_bytes = prog.has_gather ? kPrivatizeMaxBytes : kPrivatizeMaxBytesNoGather;
      if (const char* e = getenv("VKJIT_PRIV_KB")) max_bytes = std::min<size_t>((size_t)atoi(e) * 1024, kPrivatizeMaxBytesNoGather);
      const bool f32_target = ir.vars[prog.params[prog.sadd_param].var].ty == VKJIT_TY_F32;
      if (sadd_cluster() && !f32_target && bins * 4 > max_bytes) {
        // two CTAs of a cluster hold half of the privatised bins each (distributed shared memory): twice the reach,
        // half the shared memory per SM (the rest stays L1 for the trace's gathers)
        cluster_bins = true;
        kbins = (uint32_t)std::min<size_t>(bins, 2 * (kPrivatizeMaxBytesNoGather / 4));
      } else {
        kbins = (uint32_t)std::min<size_t>(bins, max_bytes / 4);
      }
    }
    // Bin-range passes (profiles/r02_h26.md, experiment 4): a target that does not fit one CTA's shared memory is
    // processed in P launches, pass p handling only the lanes whose bin lies in its range — ALL atomics of a pass go to
    // shared memory (no L2 RED at all) and, when the trace gathers with the same index, each pass only touches its slice
    // of the gather table, which then fits the L1 that is left.  Costs P reads of the index stream.  Only for traces
    // whose one side effect is that scatter_add (build_program falls back to variant 1 otherwise).
    uint32_t pass_bins = 0;
    if (kbins && !cluster_bins && scatter_passes()) {
      const size_t bins = ir.vars[prog.params[prog.sadd_param].var].array->bytes / 4;
      if (bins > kbins) {
        size_t chunk_bytes = prog.has_gather ? kPassBytesGather : kPassBytesNoGather;
        if (const char* e = getenv("VKJIT_PASS_KB")) chunk_bytes = std::min<size_t>((size_t)std::max(1, atoi(e)) * 1024, kPrivatizeMaxBytesNoGather);
        const size_t npass = (bins * 4 + chunk_bytes - 1) / chunk_bytes;
        if (npass <= kMaxPasses) pass_bins = (uint32_t)((((bins + npass - 1) / npass) + 3) & ~(size_t)3);
      }
    }
    if (!aligned || kbins) build_program(ir, roots, aligned, prog, -1, kbins ? (pass_bins ? 3 : cluster_bins ? 2 : 1) : 0);  // variant rebuild
    if (prog.privatize != 3) pass_bins = 0;

    if (trace) ts[1] = now_ns();
    CachedKernel* k = be.lookup(prog);
    if (!k) k = be.compile(ir, prog);
    if (trace) ts[2] = now_ns();

    // one fresh n*stride output per scheduled var (internal.rs:1192-1205), from the stream-ordered pool
    for (size_t r = 0; r < prog.roots.size(); ++r) outs.push_back(be.new_array((size_t)prog.n * 4));

    if (trace) ts[3] = now_ns();
    uint32_t n32 = (uint32_t)prog.n, base32 = (uint32_t)prog.base;
    ptrs.clear(); argv.clear();
    for (const Param& pr : prog.params) ptrs.push_back((uint64_t)(uintptr_t)ir.vars[pr.var].array->ptr);
    for (Array* a : outs) ptrs.push_back((uint64_t)(uintptr_t)a->ptr);
    uint32_t bin_lo = 0;
    argv.push_back(&n32); argv.push_back(&base32);
    if (prog.privatize) argv.push_back(&kbins);
    if (prog.privatize == 3) argv.push_back(&bin_lo);
    for (uint64_t& p : ptrs) argv.push_back(&p);

    const uint64_t items = prog.vectorized ? std::max<uint64_t>(prog.n >> 2, 1) : prog.n;
    if (prog.privatize) {
      // persistent CTAs of 1024 threads, each with its own copy of the privatised bins (cluster variant: each CTA
      // of a pair holds half of them; the grid is a multiple of 2)
      if (pass_bins) {
        const uint32_t bins = (uint32_t)(ir.vars[prog.params[prog.sadd_param].var].array->bytes / 4);
        for (bin_lo = 0; bin_lo < bins; bin_lo += pass_bins) {   // argv points at bin_lo / kbins
          kbins = std::min(pass_bins, bins - bin_lo);
          const uint32_t smem = kbins * 4;
          be.launch(k, (uint32_t)be.sm_count * (smem <= 100 * 1024 ? 2 : 1), 1024, argv.data(), smem);
        }
      } else if (cluster_bins) {
        const uint32_t half = ((((kbins + 1u) >> 1) + 3u) & ~3u);
        be.launch(k, (uint32_t)be.sm_count & ~1u, 1024, argv.data(), half * 4);
      } else {
        const uint32_t smem = kbins * 4;
        const uint32_t per_sm = smem <= 100 * 1024 ? 2 : 1;
        be.launch(k, (uint32_t)be.sm_count * per_sm, 1024, argv.data(), smem);
      }
    } else {
      // grid-stride launch: enough CTAs of 256 threads to fill every SM (8 x 256 = 2048 threads/SM)
      uint64_t grid = (items + 255) / 256;
      const uint64_t cap = (uint64_t)be.sm_count * 8;
      if (grid > cap) grid = cap;
      be.launch(k, (uint32_t)grid, 256, argv.data());
    }

    if (trace) ts[4] = now_ns();
    if (commit) ir.commit_roots(roots, outs, prog.order);
    else *result = outs;
    outs.clear();
    if (trace) {
      ts[5] = now_ns();
      fprintf(stderr, "[vkjit eval] nodes=%zu walk=%llu lookup=%llu alloc=%llu launch=%llu commit=%llu ns\n", prog.order.size(),
              (unsigned long long)(ts[1] - ts[0]), (unsigned long long)(ts[2] - ts[1]), (unsigned long long)(ts[3] - ts[2]),
              (unsigned long long)(ts[4] - ts[3]), (unsigned long long)(ts[5] - ts[4]));
    }
  } catch (...) {
    for (Array* a : outs) release_array(a);
    outs.clear();
    throw;
  }
}

void eval_group(Ir& ir, Backend& be, const std::vector<VarId>& roots) { run_group(ir, be, roots, true, nullptr); }

}  // namespace

// Evaluates `id` into a fresh array WITHOUT turning the var into a Binding (operands of the eager primitives
// that cannot be fused into the primitive's kernel).  A Binding is copied (the scalar variant handles
// misaligned foreign views).
Array* eval_temp(Ir& ir, VarId id) {
  std::vector<VarId> roots{id};
  std::vector<Array*> outs;
  run_group(ir, Backend::get(), roots, false, &outs);
  return outs.at(0);
}

void eval(Ir& ir, const VarId* ids, size_t n) {
  const uint64_t t0 = now_ns();
  Backend& be = Backend::get();
  ir.do_schedule(ids, n);
  if (ir.schedule.empty()) return;
  try {
    try {
      eval_group(ir, be, ir.schedule);  // the common case: one kernel for the whole schedule
    } catch (const Error& e) {
      if (e.code != VKJIT_ERR_SIZE || ir.schedule.size() < 2) throw;
      // Mixed-size schedule.  The reference asserts here (internal.rs:697-706); instead the
      // roots are evaluated in groups of equal kernel size, in schedule order.  A size conflict INSIDE one
      // root's expression still fails below.
      std::vector<VarId> pending(ir.schedule);
      Program probe;
      while (!pending.empty()) {
        std::vector<VarId> one{pending[0]};
        build_program(ir, one, true, probe);  // throws for an intrinsically inconsistent root
        const uint64_t size = probe.n;
        const bool sharded = probe.sharded;
        std::vector<VarId> group{pending[0]}, rest;
        for (size_t i = 1; i < pending.size(); ++i) {
          std::vector<VarId> r{pending[i]};
          bool same = false;
          try {
            build_program(ir, r, true, probe);
            same = probe.n == size && probe.sharded == sharded;
          } catch (const Error&) { same = false; }  // reported when its own turn comes
          (same ? group : rest).push_back(pending[i]);
        }
        eval_group(ir, be, group);
        pending.swap(rest);
      }
    }
  } catch (...) {
    ir.clear_schedule();
    throw;
  }
  ir.clear_schedule();
  g_counters.last_eval_ns = now_ns() - t0;
}

// Fused trace -> reduce: evaluates the unevaluated var `id` lane by lane and reduces it in the same
// kernel (8 B/lane for sum(x*y+c) instead of 16 B/lane when z is materialised first).  `id` stays
// unevaluated; the result is a fresh 1-element array holding this GPU's value.
Array* eval_reduce(Ir& ir, VarId id, int red) {
  Backend& be = Backend::get();
  static thread_local Program prog;
  static thread_local std::vector<void*> argv;
  static thread_local std::vector<uint64_t> ptrs;
  std::vector<VarId> roots{id};
  build_program(ir, roots, true, prog, red);
  if (prog.n == 0) fail(VKJIT_ERR_SIZE, "reduce of an empty array");
  bool aligned = true;
  for (const Param& pr : prog.params)
    if ((pr.use & USE_STREAM) && ((uintptr_t)ir.vars[pr.var].array->ptr & 15u)) aligned = false;
  if (!aligned) build_program(ir, roots, false, prog, red);
  CachedKernel* k = be.lookup(prog);
  if (!k) k = be.compile(ir, prog);
  Array* out = be.new_array(4);
  uint32_t n32 = (uint32_t)prog.n, base32 = (uint32_t)prog.base;
  ptrs.clear(); argv.clear();
  for (const Param& pr : prog.params) ptrs.push_back((uint64_t)(uintptr_t)ir.vars[pr.var].array->ptr);
  ptrs.push_back((uint64_t)(uintptr_t)be.scratch.partials);
  ptrs.push_back((uint64_t)(uintptr_t)be.scratch.ticket);
  ptrs.push_back((uint64_t)(uintptr_t)out->ptr);
  argv.push_back(&n32); argv.push_back(&base32);
  for (uint64_t& p : ptrs) argv.push_back(&p);
  const uint64_t items = prog.vectorized ? std::max<uint64_t>(prog.n >> 2, 1) : prog.n;
  uint64_t grid = (items + 255) / 256;
  const uint64_t cap = std::min<uint64_t>((uint64_t)be.sm_count * 8, prims::kReduceMaxCtas);
  if (grid > cap) grid = cap;
  try {
    be.launch(k, (uint32_t)grid, 256, argv.data());
  } catch (...) { release_array(out); throw; }
  return out;
}

// Fused trace -> prefix-sum / compress (scan_fused.cuh): root 0 is scanned (or is the compress mask), root 1
// (SCAN_COMPRESS_VALUE) supplies the compacted values.  None of the roots is materialised or committed.
// Returns false when this trace is not fused (side effects, > 6 streamed arrays, a misaligned streamed
// array, or a body so large that evaluating it twice / inlining it 24 times does not pay): the caller then
// evaluates the operands into temporaries and runs the hand-written primitive.
bool eval_scan(Ir& ir, int mode, const std::vector<VarId>& roots, const uint32_t* initial, uint32_t* count_dev, Array** out,
               uint64_t* n_out, const uint32_t* index_base) {
  Backend& be = Backend::get();
  static thread_local Program prog;
  build_program(ir, roots, true, prog, -1, false, mode);
  *n_out = prog.n;
  if (!prog.vectorized) return false;  // > kMaxVectorNodes nodes
  if (getenv("VKJIT_NO_FUSED_SCAN")) return false;
  size_t ns = 0;
  for (const Param& pr : prog.params) {
    if (pr.use & USE_SCATTER) return false;
    if (pr.use & USE_STREAM) {
      ++ns;
      if ((uintptr_t)ir.vars[pr.var].array->ptr & 15u) return false;
    }
  }
  if (ns > (size_t)kScanFusedMaxStreams) return false;
  if (prog.n == 0) { *out = be.new_array(0); return true; }  // nothing to scan: empty result, no launch
  CachedKernel* k = be.lookup(prog);
  if (!k) k = be.compile(ir, prog);

  const ScanFusedGeom geom = scan_fused_geom(ns, mode, prog.order.size());
  const size_t tile = geom.tile();
  const size_t tiles = (prog.n + tile - 1) / tile;
  // control-warp kernels keep a packed copy of the tile aggregates (8 bytes per tile) behind the status lines
  const bool has_packed = geom.ctrl || (geom.lag && scan_lag_packed());
  const size_t packed_words = has_packed ? tiles + 1024 : 0;
  be.ensure_scan_scratch(prog.n + (has_packed ? prog.n / 8 + 128 * tile : 0), tile);
  // status lines: a header line + one per tile (what ensure_scan_scratch sizes); control-warp kernels: one more line, then the packed array
  const size_t clear_words = has_packed ? (size_t)prims::kStatusWordsPerTile * (2 + tiles) + packed_words : (size_t)prims::kStatusWordsPerTile * (1 + tiles);
  if (clear_words > be.scratch.tile_state_words) fail(VKJIT_ERR_INVALID, "scan scratch too small");
  ck(cudaMemsetAsync(be.scratch.tile_state, 0, clear_words * 8, (cudaStream_t)be.enqueue_stream()), "scan status memset");
  Array* o = be.new_array((size_t)prog.n * 4);  // compress: worst case, trimmed by the caller
  // VkPtrs: streamed arrays (at least one slot), then the gather pointers in parameter order
  std::vector<uint64_t> block;
  for (const Param& pr : prog.params) if (pr.use & USE_STREAM) block.push_back((uint64_t)(uintptr_t)ir.vars[pr.var].array->ptr);
  if (block.empty()) block.push_back(0);
  for (const Param& pr : prog.params) if (pr.use & (USE_GATHER | USE_SCATTER)) block.push_back((uint64_t)(uintptr_t)ir.vars[pr.var].array->ptr);
  uint32_t n32 = (uint32_t)prog.n, base32 = (uint32_t)prog.base, tiles32 = (uint32_t)tiles;
  uint64_t outp = (uint64_t)(uintptr_t)o->ptr, cntp = (uint64_t)(uintptr_t)count_dev, statep = (uint64_t)(uintptr_t)be.scratch.tile_state,
           initp = (uint64_t)(uintptr_t)initial, ibasep = (uint64_t)(uintptr_t)index_base;
  void* argv[] = {&n32, &base32, block.data(), &outp, &cntp, &tiles32, &statep, &initp, &ibasep};
  try {
    be.launch(k, (uint32_t)std::min<size_t>(tiles, (size_t)be.sm_count * k->ctas_per_sm), (uint32_t)geom.launch_threads(), argv, (uint32_t)geom.smem(ns));
  } catch (...) { release_array(o); throw; }
  if (const char* tf = fscan_trace_file()) {  // per-tile phase stamps live in the spare words of the status lines
    ck(cudaStreamSynchronize((cudaStream_t)be.enqueue_stream()), "sync");
    std::vector<uint64_t> h((size_t)prims::kStatusWordsPerTile * (1 + tiles));
    ck(cudaMemcpy(h.data(), be.scratch.tile_state, h.size() * 8, cudaMemcpyDeviceToHost), "trace readback");
    if (FILE* fp = fopen(tf, "wb")) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
  }
  *out = o;
  return true;
}

}  // namespace vkjit
