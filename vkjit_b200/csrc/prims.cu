// prims.cu — hand-written CUDA primitives for sm_100a: horizontal reductions, the peer-memory exchanges of the
// multi-GPU path and small helpers (prefix-sum / stream compaction live in scan.cu).  All of them are HBM-bound integer /
// byte movers: the design rules that matter are coalesced 128-bit accesses, enough loads in
// flight per SM to cover HBM latency, and grids sized in multiples of the SM count
// (148 on B200).  No tensor cores on purpose.
//
// There is no reference implementation of these operations; the CPU oracle
// (oracle/oracle.cpp) is the specification they are tested against.
#include "prims.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "common.h"

namespace vkjit {
namespace prims {

// ---------------------------------------------------------------------------------------
// helpers
// ---------------------------------------------------------------------------------------
#include "scan_common.cuh"  // streaming ld/st, tile status words + look-back, TMA/mbarrier (shared with the NVRTC kernels)

template <typename T> __device__ __forceinline__ T from_bits(uint32_t w);
template <> __device__ __forceinline__ uint32_t from_bits<uint32_t>(uint32_t w) { return w; }
template <> __device__ __forceinline__ int32_t from_bits<int32_t>(uint32_t w) { return (int32_t)w; }
template <> __device__ __forceinline__ float from_bits<float>(uint32_t w) { return __uint_as_float(w); }
__device__ __forceinline__ uint32_t to_bits(uint32_t v) { return v; }
__device__ __forceinline__ uint32_t to_bits(int32_t v) { return (uint32_t)v; }
__device__ __forceinline__ uint32_t to_bits(float v) { return __float_as_uint(v); }

template <typename T, int RED> struct RedOp;
template <typename T> struct RedOp<T, VKJIT_RED_SUM> {
  __device__ static T identity() { return T(0); }
  __device__ static T apply(T a, T b) { return a + b; }  // u32/i32 wrap mod 2^32; f32 any association
};
template <> struct RedOp<int32_t, VKJIT_RED_SUM> {
  __device__ static int32_t identity() { return 0; }
  __device__ static int32_t apply(int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); }
};
template <> struct RedOp<uint32_t, VKJIT_RED_MIN> {
  __device__ static uint32_t identity() { return 0xFFFFFFFFu; }
  __device__ static uint32_t apply(uint32_t a, uint32_t b) { return min(a, b); }
};
template <> struct RedOp<uint32_t, VKJIT_RED_MAX> {
  __device__ static uint32_t identity() { return 0u; }
  __device__ static uint32_t apply(uint32_t a, uint32_t b) { return max(a, b); }
};
template <> struct RedOp<int32_t, VKJIT_RED_MIN> {
  __device__ static int32_t identity() { return 0x7FFFFFFF; }
  __device__ static int32_t apply(int32_t a, int32_t b) { return min(a, b); }
};
template <> struct RedOp<int32_t, VKJIT_RED_MAX> {
  __device__ static int32_t identity() { return (int32_t)0x80000000; }
  __device__ static int32_t apply(int32_t a, int32_t b) { return max(a, b); }
};
// fminf/fmaxf: NaN-ignoring, -0 < +0 (PTX min/max.f32) — matches the oracle's fmin_spec/fmax_spec.
// Identity = NaN so that it is ignored by every real operand.
template <> struct RedOp<float, VKJIT_RED_MIN> {
  __device__ static float identity() { return __uint_as_float(0x7FC00000u); }
  __device__ static float apply(float a, float b) { return fminf(a, b); }
};
template <> struct RedOp<float, VKJIT_RED_MAX> {
  __device__ static float identity() { return __uint_as_float(0x7FC00000u); }
  __device__ static float apply(float a, float b) { return fmaxf(a, b); }
};

template <typename T, int RED>
__device__ __forceinline__ T warp_reduce(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = RedOp<T, RED>::apply(v, from_bits<T>(__shfl_xor_sync(0xFFFFFFFFu, to_bits(v), o)));
  return v;
}

// Block-wide reduction: warp shuffle, then a shared-memory tree over the warp results.
template <typename T, int RED, int THREADS>
__device__ __forceinline__ T block_reduce(T v, T* smem /* THREADS/32 */) {
  constexpr int WARPS = THREADS / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_reduce<T, RED>(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    T w = lane < WARPS ? smem[lane] : RedOp<T, RED>::identity();
    w = warp_reduce<T, RED>(w);
    if (lane == 0) smem[0] = w;
  }
  __syncthreads();
  T r = smem[0];
  __syncthreads();
  return r;
}

// ---------------------------------------------------------------------------------------
// one-shot all-reduce over NVLink peer memory (called by ONE thread of ONE CTA per GPU)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void sys_store(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t sys_load(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// Waits until rank r's word of exchange `seq` has arrived in this GPU's mailbox.  A peer that never arrives (crashed
// process, rank stuck in host code for longer than the watchdog) does not hang the GPU: the exchange gives up, flags
// the fault in host-mapped memory — the next synchronising call of the backend returns VKJIT_ERR_DIST — and the
// kernel finishes with an unspecified value.
__device__ __forceinline__ uint64_t mailbox_wait(const uint64_t* word, const Mailbox& mb, const unsigned long long t0) {
  uint64_t w = sys_load(word);
  uint32_t spins = 0;
  while ((uint32_t)(w >> 32) != mb.seq) {
    if ((++spins & 1023u) == 0 && mb.timeout_ns && global_ns() - t0 > mb.timeout_ns) {
      if (mb.fault) { *(volatile uint32_t*)mb.fault = mb.seq; __threadfence_system(); }
      break;
    }
    w = sys_load(word);
  }
  return w;
}

template <typename T, int RED>
__device__ __forceinline__ T mailbox_allreduce(T mine, const Mailbox& mb) {
  const size_t slot = (size_t)(mb.seq % kMailSlots) * kMailRanks;
  const uint64_t word = ((uint64_t)mb.seq << 32) | to_bits(mine);
  for (int p = 0; p < mb.world; ++p) sys_store(mb.peers[p] + slot + mb.rank, word);  // P2P stores over NVLink
  const uint64_t* local = mb.local + slot;
  T acc = RedOp<T, RED>::identity();
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < mb.world; ++r)  // fixed rank order: same bits on every GPU
    acc = RedOp<T, RED>::apply(acc, from_bits<T>((uint32_t)mailbox_wait(local + r, mb, t0)));
  return acc;
}

// Exclusive scan over ranks of one u32 per GPU (mod 2^32): out[0] = sum of `mine` over all ranks below this one;
// TOTAL: additionally out[1] = sum over all ranks.  EVERY exchange waits for every rank, also the plain exscan that
// only needs the lower ranks' words: a rank that waited for nobody could run arbitrarily far ahead of the others and
// rewrite ring slot seq % kMailSlots before a slower rank has read it.  With all ranks waiting for all, a rank can
// publish exchange s + 1 only after every peer has published s, i.e. has finished reading s - 1: no rank is ever
// more than one exchange ahead of another and the 64-slot ring cannot wrap.
template <bool TOTAL>
__global__ void p2p_exscan_kernel(const uint32_t* mine, uint32_t* out, Mailbox mb) {
  if (threadIdx.x != 0) return;
  const size_t slot = (size_t)(mb.seq % kMailSlots) * kMailRanks;
  const uint64_t word = ((uint64_t)mb.seq << 32) | mine[0];
  for (int p = 0; p < mb.world; ++p) sys_store(mb.peers[p] + slot + mb.rank, word);
  const uint64_t* local = mb.local + slot;
  uint32_t acc = 0, below = 0;
  const unsigned long long t0 = global_ns();
  for (int r = 0; r < mb.world; ++r) {
    const uint64_t w = mailbox_wait(local + r, mb, t0);
    if (r == mb.rank) below = acc;
    acc += (uint32_t)w;
  }
  out[0] = below;
  if (TOTAL) out[1] = acc;
}

// out[0] = sum of v[0..rank) — the NCCL path all-reduces a one-hot vector of the per-rank totals first
__global__ void prefix_of_rank_kernel(const uint32_t* v, int rank, uint32_t* out) {
  if (threadIdx.x != 0) return;
  uint32_t acc = 0;
  for (int r = 0; r < rank; ++r) acc += v[r];
  out[0] = acc;
}

__global__ void one_hot_kernel(const uint32_t* mine, int rank, int world, uint32_t* v) {
  if ((int)threadIdx.x < world) v[threadIdx.x] = (int)threadIdx.x == rank ? mine[0] : 0u;
}

template <typename T, int RED>
__global__ void p2p_allreduce_kernel(uint32_t* out, Mailbox mb) {
  if (threadIdx.x == 0) out[0] = to_bits(mailbox_allreduce<T, RED>(from_bits<T>(out[0]), mb));
}

// ---------------------------------------------------------------------------------------
// horizontal reduction
// ---------------------------------------------------------------------------------------
// Algorithmic traffic: 4 B/lane read, 4 B written in total.
//
// Programmatic dependent launch (PDL): the kernel is launched with programmatic stream serialization, so its CTAs may
// start while the PREVIOUS kernel on the stream is still in its tail.  A reduction has a long streaming phase that
// only reads its input and a short tail that touches state shared between launches (partials, ticket, result,
// mailbox).  Each CTA therefore (1) streams, (2) signals `launch_dependents` — once every CTA has, the next
// reduction's CTAs take over the SM slots this grid frees — and (3) executes `griddepcontrol.wait` (completion and
// memory flush of the previous kernel) before its tail.  The last-CTA fold and, on several GPUs, the NVLink exchange
// of reduction k thus run underneath the streaming phase of reduction k+1, while the tails stay strictly ordered.
// flags bit 0 (kReduceWaitFirst): the input may have been written by a kernel that is still running (the host cannot
// prove otherwise, see capi.cpp: vkjit_reduce) — wait before the first load, i.e. plain stream order.
template <typename T, int RED, int THREADS>
__global__ void __launch_bounds__(THREADS, 2048 / THREADS)
reduce_kernel(const uint32_t* __restrict__ in, size_t n, uint32_t* __restrict__ partials, unsigned int* __restrict__ ticket,
              uint32_t* __restrict__ out, const Mailbox mb, const uint32_t flags, unsigned long long* __restrict__ trace) {
  using O = RedOp<T, RED>;
  __shared__ T smem[THREADS / 32];
  __shared__ bool is_last;
  if (flags & kReduceWaitFirst) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (trace && blockIdx.x == 0 && threadIdx.x == 0) trace[0] = global_ns();

  const size_t tid = (size_t)blockIdx.x * THREADS + threadIdx.x;
  const size_t nthreads = (size_t)gridDim.x * THREADS;
  const size_t n4 = n >> 2;
  const uint4* in4 = reinterpret_cast<const uint4*>(in);

  T a0 = O::identity(), a1 = O::identity(), a2 = O::identity(), a3 = O::identity();
  size_t v = tid;
  // 4 independent 128-bit loads in flight per thread (64 B/thread, 128 KB/SM at full occupancy)
  for (; v + 3 * nthreads < n4; v += 4 * nthreads) {
    const uint4 x0 = ld_stream(in4 + v);
    const uint4 x1 = ld_stream(in4 + v + nthreads);
    const uint4 x2 = ld_stream(in4 + v + 2 * nthreads);
    const uint4 x3 = ld_stream(in4 + v + 3 * nthreads);
    a0 = O::apply(a0, from_bits<T>(x0.x)); a1 = O::apply(a1, from_bits<T>(x0.y));
    a2 = O::apply(a2, from_bits<T>(x0.z)); a3 = O::apply(a3, from_bits<T>(x0.w));
    a0 = O::apply(a0, from_bits<T>(x1.x)); a1 = O::apply(a1, from_bits<T>(x1.y));
    a2 = O::apply(a2, from_bits<T>(x1.z)); a3 = O::apply(a3, from_bits<T>(x1.w));
    a0 = O::apply(a0, from_bits<T>(x2.x)); a1 = O::apply(a1, from_bits<T>(x2.y));
    a2 = O::apply(a2, from_bits<T>(x2.z)); a3 = O::apply(a3, from_bits<T>(x2.w));
    a0 = O::apply(a0, from_bits<T>(x3.x)); a1 = O::apply(a1, from_bits<T>(x3.y));
    a2 = O::apply(a2, from_bits<T>(x3.z)); a3 = O::apply(a3, from_bits<T>(x3.w));
  }
  for (; v < n4; v += nthreads) {
    const uint4 x0 = ld_stream(in4 + v);
    a0 = O::apply(a0, from_bits<T>(x0.x)); a1 = O::apply(a1, from_bits<T>(x0.y));
    a2 = O::apply(a2, from_bits<T>(x0.z)); a3 = O::apply(a3, from_bits<T>(x0.w));
  }
  for (size_t i = (n4 << 2) + tid; i < n; i += nthreads) a0 = O::apply(a0, from_bits<T>(in[i]));

  asm volatile("griddepcontrol.launch_dependents;");  // this CTA's input reads are done
  T acc = O::apply(O::apply(a0, a1), O::apply(a2, a3));
  acc = block_reduce<T, RED, THREADS>(acc, smem);
  unsigned long long ts1 = 0, ts2 = 0;
  if (trace && threadIdx.x == 0) ts1 = global_ns();
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous kernel on the stream has completed (no-op if it already had)
  if (trace && threadIdx.x == 0) ts2 = global_ns();

  // publish the CTA partial; the last CTA to arrive folds all partials in a fixed order.  One acquire-release
  // atomic on the ticket orders the partial store before it and the partial loads of the last CTA after it
  // (no separate __threadfence round trips on the critical path of small launches).
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = to_bits(acc);
    unsigned int t;
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    if (trace && threadIdx.x == 0) { trace[1] = ts1; trace[2] = ts2; trace[3] = global_ns(); }
    T f = O::identity();
    for (unsigned int i = threadIdx.x; i < gridDim.x; i += THREADS) f = O::apply(f, from_bits<T>(__ldcg(partials + i)));
    f = block_reduce<T, RED, THREADS>(f, smem);
    if (threadIdx.x == 0) {
      *ticket = 0u;  // self-reset for the next launch on this stream
      if (trace) trace[4] = global_ns();
      // fused collective: the per-GPU partial goes straight to every peer's mailbox (no second
      // kernel, no NCCL launch); world == 1 compiles to the plain store
      if (mb.world > 1) f = mailbox_allreduce<T, RED>(f, mb);
      out[0] = to_bits(f);
      if (trace) { trace[5] = global_ns(); trace[7] = (unsigned long long)mb.world; }
    }
  }
}

// $VKJIT_REDUCE_TRACE ring (see prims.h)
static unsigned long long* g_reduce_trace = nullptr;
static size_t g_reduce_trace_launches = 0;
static unsigned long long* next_reduce_trace(cudaStream_t s) {
  static const bool on = getenv("VKJIT_REDUCE_TRACE") != nullptr;
  if (!on) return nullptr;
  if (!g_reduce_trace) {
    const size_t bytes = (size_t)kReduceTraceLaunches * kReduceTraceWords * 8;
    if (cudaMalloc(&g_reduce_trace, bytes) != cudaSuccess || cudaMemset(g_reduce_trace, 0, bytes) != cudaSuccess)
      fail(VKJIT_ERR_CUDA, "reduce trace allocation failed");
  }
  unsigned long long* slot = g_reduce_trace + (g_reduce_trace_launches % kReduceTraceLaunches) * kReduceTraceWords;
  // no per-launch memset: it would sit between the kernels on the stream and break the programmatic overlap that is
  // being measured; every launch overwrites all stamps of its slot
  (void)s;
  ++g_reduce_trace_launches;
  return slot;
}

size_t reduce_trace_dump(unsigned long long* out, size_t cap_words, void* stream) {
  if (!g_reduce_trace) return 0;
  cudaStreamSynchronize((cudaStream_t)stream);
  const size_t launches = std::min<size_t>(g_reduce_trace_launches, kReduceTraceLaunches);
  const size_t words = std::min(cap_words, launches * kReduceTraceWords);
  // oldest launch first
  const size_t first = g_reduce_trace_launches > (size_t)kReduceTraceLaunches ? g_reduce_trace_launches % kReduceTraceLaunches : 0;
  for (size_t i = 0; i * kReduceTraceWords < words; ++i) {
    const size_t src = (first + i) % kReduceTraceLaunches;
    cudaMemcpy(out + i * kReduceTraceWords, g_reduce_trace + src * kReduceTraceWords, kReduceTraceWords * 8, cudaMemcpyDeviceToHost);
    out[i * kReduceTraceWords + 6] = g_reduce_trace_launches - launches + i;
  }
  const size_t n = words / kReduceTraceWords;
  g_reduce_trace_launches = 0;
  return n;
}

template <typename T, int RED>
static void launch_reduce(const void* in, size_t n, void* out, const Scratch& sc, int sm_count, cudaStream_t s, const Mailbox& mb,
                          uint32_t flags) {
  const size_t n4 = n >> 2;
  size_t ctas = (std::max<size_t>(n4, 1) + kReduceThreads - 1) / kReduceThreads;
  // One wave MINUS TWO CTAs (4 CTAs of 512 threads per SM).  Back-to-back reductions hand the SMs over through
  // programmatic dependent launch: the CTAs of reduction k+1 take the slots the CTAs of k free when they exit.  All of
  // them exit right after the ticket — except the one that folds (and, on several GPUs, sits in the peer exchange for
  // ~10 us).  With a grid of exactly one wave, ONE CTA of k+1 finds no slot until that tail is over, starts 7 us (one
  // GPU) to 17 us (eight GPUs) late and then streams its fixed 1/592 share almost alone: measured launch-to-launch
  // period 22.0 / 24.5 us for a 128 MiB shard that streams in 19.5 (profiles/r02_n8_timeline.md).  Two spare slots
  // (a second tail can still be alive when ranks are skewed) cost 0.34 % of the streaming parallelism.
  const size_t wave = (size_t)sm_count * (2048 / kReduceThreads);
  static const int spare = [] { const char* e = getenv("VKJIT_REDUCE_SPARE_CTAS"); return e ? atoi(e) : 2; }();
  const size_t cap = wave > 8 + (size_t)spare ? wave - (size_t)spare : wave;
  if (ctas > cap) ctas = cap;
  if (ctas > (size_t)kReduceMaxCtas) ctas = kReduceMaxCtas;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)ctas);
  cfg.blockDim = dim3(kReduceThreads);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  unsigned long long* trace = next_reduce_trace(s);
  const cudaError_t e = cudaLaunchKernelEx(&cfg, reduce_kernel<T, RED, kReduceThreads>, (const uint32_t*)in, n, (uint32_t*)sc.partials,
                                           sc.ticket, (uint32_t*)out, mb, flags, trace);
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("reduce launch: ") + cudaGetErrorString(e));
}

void reduce(int red, uint32_t ty, const void* in, size_t n, void* out, const Scratch& sc, int sm_count, void* stream,
            const Mailbox* mailbox, uint32_t flags) {
  cudaStream_t s = (cudaStream_t)stream;
  Mailbox mb;
  if (mailbox) mb = *mailbox;
#define VK_DISPATCH(T)                                                                          \
  switch (red) {                                                                                \
    case VKJIT_RED_SUM: launch_reduce<T, VKJIT_RED_SUM>(in, n, out, sc, sm_count, s, mb, flags); break;    \
    case VKJIT_RED_MIN: launch_reduce<T, VKJIT_RED_MIN>(in, n, out, sc, sm_count, s, mb, flags); break;    \
    case VKJIT_RED_MAX: launch_reduce<T, VKJIT_RED_MAX>(in, n, out, sc, sm_count, s, mb, flags); break;    \
    default: fail(VKJIT_ERR_INVALID, "unknown reduction");                                      \
  }
  switch (ty) {
    case VKJIT_TY_U32: VK_DISPATCH(uint32_t) break;
    case VKJIT_TY_I32: VK_DISPATCH(int32_t) break;
    case VKJIT_TY_F32: VK_DISPATCH(float) break;
    default: fail(VKJIT_ERR_TYPE, "reduce needs U32/I32/F32");
  }
#undef VK_DISPATCH
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("reduce launch: ") + cudaGetErrorString(e));
}

void p2p_allreduce(int red, uint32_t ty, void* out, const Mailbox& mb, void* stream) {
  cudaStream_t s = (cudaStream_t)stream;
#define VK_P2P(T)                                                                                             \
  switch (red) {                                                                                              \
    case VKJIT_RED_SUM: p2p_allreduce_kernel<T, VKJIT_RED_SUM><<<1, 32, 0, s>>>((uint32_t*)out, mb); break;   \
    case VKJIT_RED_MIN: p2p_allreduce_kernel<T, VKJIT_RED_MIN><<<1, 32, 0, s>>>((uint32_t*)out, mb); break;   \
    case VKJIT_RED_MAX: p2p_allreduce_kernel<T, VKJIT_RED_MAX><<<1, 32, 0, s>>>((uint32_t*)out, mb); break;   \
    default: fail(VKJIT_ERR_INVALID, "unknown reduction");                                                    \
  }
  switch (ty) {
    case VKJIT_TY_U32: VK_P2P(uint32_t) break;
    case VKJIT_TY_I32: VK_P2P(int32_t) break;
    case VKJIT_TY_F32: VK_P2P(float) break;
    default: fail(VKJIT_ERR_TYPE, "reduce needs U32/I32/F32");
  }
#undef VK_P2P
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("p2p all-reduce launch: ") + cudaGetErrorString(e));
}

void p2p_exscan_u32(const uint32_t* mine, uint32_t* out, const Mailbox& mb, void* stream) {
  p2p_exscan_kernel<false><<<1, 32, 0, (cudaStream_t)stream>>>(mine, out, mb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("p2p exscan launch: ") + cudaGetErrorString(e));
}
void p2p_exscan_total_u32(const uint32_t* mine, uint32_t* out2, const Mailbox& mb, void* stream) {
  p2p_exscan_kernel<true><<<1, 32, 0, (cudaStream_t)stream>>>(mine, out2, mb);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("p2p exscan launch: ") + cudaGetErrorString(e));
}
void one_hot_u32(const uint32_t* mine, int rank, int world, uint32_t* v, void* stream) {
  one_hot_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(mine, rank, world, v);
}
void prefix_of_rank_u32(const uint32_t* v, int rank, uint32_t* out, void* stream) {
  prefix_of_rank_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(v, rank, out);
}

__global__ void fill_kernel(uint32_t* out, uint32_t value, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) out[i] = value;
}
void fill_u32(uint32_t* out, uint32_t value, size_t n, void* stream) {
  if (n == 0) return;
  const unsigned blocks = (unsigned)std::min<size_t>((n + 255) / 256, 148 * 8);
  fill_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(out, value, n);
}

}  // namespace prims
}  // namespace vkjit
