// dist.cpp — see dist.h.  NCCL is bound with dlopen so that single-GPU use (and loading the
// library on a machine without NCCL) needs nothing; inside a torch process the already loaded
// libnccl.so.2 is reused.
#include "dist.h"

#include <dlfcn.h>
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cstdlib>

#include <string>
#include <vector>

#include "runtime.h"

namespace vkjit {
namespace dist {
namespace {

struct Nccl {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  void load() {
    if (handle) return;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) fail(VKJIT_ERR_DIST, std::string("cannot load libnccl.so.2: ") + dlerror());
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(handle, "ncclAllReduce");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce || !GetErrorString)
      fail(VKJIT_ERR_DIST, "libnccl.so.2 lacks required entry points");
  }
};

Nccl g_nccl;
ncclComm_t g_comm = nullptr;
int g_rank = 0, g_world = 1;
bool g_active = false;

// peer mailboxes
uint64_t* g_mail_local = nullptr;
uint64_t* g_mail_peer[prims::kMailRanks] = {};
uint64_t** g_mail_table = nullptr;  // device copy of g_mail_peer
bool g_p2p = false;
uint32_t g_seq = 0;
uint32_t* g_fault_host = nullptr;  // host-mapped word the exchange watchdog writes (prims.cu: mailbox_wait)
uint32_t* g_fault_dev = nullptr;
uint64_t g_timeout_ns = 0;

void ckn(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) fail(VKJIT_ERR_DIST, std::string(what) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "nccl error"));
}

}  // namespace

bool active() { return g_active; }
int rank() { return g_rank; }
int world() { return g_world; }

void unique_id(void* out128) {
  g_nccl.load();
  ncclUniqueId id;
  ckn(g_nccl.GetUniqueId(&id), "ncclGetUniqueId");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  memcpy(out128, &id, 128);
}

void init(int rank, int world, const void* id128) {
  if (world < 1 || rank < 0 || rank >= world) fail(VKJIT_ERR_INVALID, "bad rank/world");
  if (g_active) fail(VKJIT_ERR_DIST, "vkjit_dist_init called twice");
  Backend::get();  // needs the device bound first
  g_rank = rank; g_world = world;
  if (world > 1) {
    g_nccl.load();
    ncclUniqueId id;
    memcpy(&id, id128, 128);
    ckn(g_nccl.CommInitRank(&g_comm, world, id, rank), "ncclCommInitRank");
  }
  g_active = true;
}

void init_env() {
  if (g_active) fail(VKJIT_ERR_DIST, "vkjit_dist_init_env: already initialised");
  auto env_int = [](const char* n, int d) { const char* v = getenv(n); return (v && *v) ? atoi(v) : d; };
  const int world = env_int("WORLD_SIZE", 1), rank = env_int("RANK", 0);
  if (!Backend::initialized()) Backend::init(env_int("LOCAL_RANK", rank));
  unsigned char id[128] = {0}, mine[64] = {0};
  std::vector<unsigned char> all((size_t)std::max(world, 1) * 64);
  if (world > 1) {
    if (rank == 0) unique_id(id);
    mailbox_handle(mine);
    std::string addr; int port = 0;
    rendezvous_endpoint(addr, port);
    const char* to = getenv("VKJIT_RDZV_TIMEOUT_S");
    rendezvous(rank, world, addr.c_str(), port, mine, 64, id, 128, all.data(), to ? atof(to) : 120.0);
  }
  init(rank, world, id);
  if (world > 1) mailbox_open(all.data(), world);
}

void mailbox_handle(void* out64) {
  Backend::get();
  if (!g_mail_local) {
    void* p = nullptr;
    const size_t bytes = (size_t)prims::kMailSlots * prims::kMailRanks * 8;
    if (cudaMalloc(&p, bytes) != cudaSuccess || cudaMemset(p, 0, bytes) != cudaSuccess) fail(VKJIT_ERR_CUDA, "mailbox allocation failed");
    g_mail_local = (uint64_t*)p;
  }
  cudaIpcMemHandle_t hnd;
  cudaError_t e = cudaIpcGetMemHandle(&hnd, g_mail_local);
  if (e != cudaSuccess) fail(VKJIT_ERR_DIST, std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
  memcpy(out64, &hnd, 64);
}

void mailbox_open(const void* handles, int world) {
  if (!g_active || world != g_world) fail(VKJIT_ERR_DIST, "mailbox_open: call vkjit_dist_init first with the same world size");
  if (world > prims::kMailRanks) fail(VKJIT_ERR_DIST, "mailbox supports at most 8 ranks");
  if (!g_mail_local) fail(VKJIT_ERR_DIST, "mailbox_open before mailbox_handle");
  for (int r = 0; r < world; ++r) {
    if (r == g_rank) { g_mail_peer[r] = g_mail_local; continue; }
    cudaIpcMemHandle_t hnd;
    memcpy(&hnd, (const char*)handles + (size_t)r * 64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) fail(VKJIT_ERR_DIST, std::string("cudaIpcOpenMemHandle (rank ") + std::to_string(r) + "): " + cudaGetErrorString(e));
    g_mail_peer[r] = (uint64_t*)p;
  }
  if (!g_mail_table) {
    void* t = nullptr;
    if (cudaMalloc(&t, sizeof(g_mail_peer)) != cudaSuccess) fail(VKJIT_ERR_CUDA, "mailbox table allocation failed");
    g_mail_table = (uint64_t**)t;
  }
  if (cudaMemcpy(g_mail_table, g_mail_peer, sizeof(g_mail_peer), cudaMemcpyHostToDevice) != cudaSuccess)
    fail(VKJIT_ERR_CUDA, "mailbox table upload failed");
  if (!g_fault_host) {
    void* hp = nullptr; void* dp = nullptr;
    if (cudaHostAlloc(&hp, 64, cudaHostAllocMapped) != cudaSuccess || cudaHostGetDevicePointer(&dp, hp, 0) != cudaSuccess)
      fail(VKJIT_ERR_CUDA, "fault word allocation failed");
    memset(hp, 0, 64);
    g_fault_host = (uint32_t*)hp; g_fault_dev = (uint32_t*)dp;
    set_fault_word(g_fault_host);
  }
  // NCCL has no deadline at all; the watchdog only exists so that a crashed peer cannot hang this GPU forever
  const char* to = getenv("VKJIT_DIST_TIMEOUT_S");
  const double secs = to ? atof(to) : 120.0;
  g_timeout_ns = secs > 0 ? (uint64_t)(secs * 1e9) : 0;
  const char* force = getenv("VKJIT_DIST");
  g_p2p = !(force && std::string(force) == "nccl");
  g_seq = 0;
}

bool p2p_enabled() { return g_active && g_world > 1 && g_p2p; }

void set_p2p(bool on) {
  if (on && !g_mail_peer[g_rank]) fail(VKJIT_ERR_DIST, "mailboxes are not open");
  if (Backend::initialized()) Backend::get().sync();
  g_p2p = on;
}

prims::Mailbox next_mailbox() {
  prims::Mailbox mb;
  mb.peers = g_mail_table; mb.local = g_mail_local;
  mb.rank = g_rank; mb.world = g_world;
  mb.seq = ++g_seq;
  if (g_seq == 0xFFFFFFFFu) g_seq = 0;
  mb.fault = g_fault_dev; mb.timeout_ns = g_timeout_ns;
  Backend::counters().collectives += 1;  // no stream operation of its own: the exchange runs inside the reduce kernel
  return mb;
}

void shutdown() {
  if (g_p2p || g_mail_local) {
    if (Backend::initialized()) Backend::get().sync();
    for (int r = 0; r < prims::kMailRanks; ++r) {
      if (g_mail_peer[r] && g_mail_peer[r] != g_mail_local) cudaIpcCloseMemHandle(g_mail_peer[r]);
      g_mail_peer[r] = nullptr;
    }
    if (g_mail_local) { cudaFree(g_mail_local); g_mail_local = nullptr; }
    if (g_mail_table) { cudaFree(g_mail_table); g_mail_table = nullptr; }
    if (g_fault_host) { set_fault_word(nullptr); cudaFreeHost(g_fault_host); g_fault_host = g_fault_dev = nullptr; }
    g_p2p = false; g_seq = 0;
  }
  if (g_comm) { g_nccl.CommDestroy(g_comm); g_comm = nullptr; }
  g_active = false; g_rank = 0; g_world = 1;
}

void allreduce(void* buf, uint32_t ty, int red, size_t count) {
  if (!g_active || g_world == 1) return;
  ncclDataType_t dt = ty == VKJIT_TY_F32 ? ncclFloat32 : ty == VKJIT_TY_I32 ? ncclInt32 : ncclUint32;
  ncclRedOp_t op = red == VKJIT_RED_SUM ? ncclSum : red == VKJIT_RED_MIN ? ncclMin : ncclMax;
  ckn(g_nccl.AllReduce(buf, buf, count, dt, op, g_comm, (cudaStream_t)Backend::get().enqueue_stream()), "ncclAllReduce");
  Backend::counters().collectives += 1;
  Backend::counters().stream_ops += 1;
}

void shard_range(size_t n, int rank, int world, size_t& lo, size_t& hi) {
  if (world < 1 || rank < 0 || rank >= world) fail(VKJIT_ERR_INVALID, "bad rank/world");
  const size_t q = (n / 4) / (size_t)world * 4;
  lo = q * (size_t)rank;
  hi = rank == world - 1 ? n : q * (size_t)(rank + 1);
}

}  // namespace dist
}  // namespace vkjit
