// scan_common.cuh — device helpers shared by the hand-written scan/compress kernels (scan.cu; prims.cu uses the ld/st helpers — compiled by
// nvcc) and the generated fused trace -> scan kernels (scan_fused.cuh, compiled by NVRTC at run time; the
// Makefile embeds both texts into the library as strings).  No #includes: only built-in types and intrinsics.
// uint32_t / uint64_t are typedef'd by the NVRTC prelude.
// Streaming 128-bit load: read-only path, do not allocate in L1 (every byte is used once).
__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}

enum : uint32_t { ST_INVALID = 0, ST_AGGREGATE = 1, ST_INCLUSIVE = 2 };

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t warp_sum(uint32_t v) { return __reduce_add_sync(0xFFFFFFFFu, v); }

// Status words are read and written with gpu-scope relaxed accesses (L2 is the coherence point;
// `volatile` would compile to system-scope STRONG.SYS) and every tile owns a full 128-byte line:
// a window of predecessors then maps to many L2 slices instead of hammering the one or two
// slices that hold a packed status array (the packed layout made every poll round take ~0.7 us
// and the nearest window was polled 6.6 times per tile — ncu source page, profiles/).
constexpr int kStatusStride = 16;  // 64-bit words per tile status slot (128 bytes)
__device__ __forceinline__ uint64_t status_load(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void status_store(uint64_t* p, uint64_t v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// Executed by warp 0 of the CTA owning `tile`.  Returns the exclusive prefix of the tile.
// One round inspects kLookWide x 32 = 160 predecessors with all status loads in flight together:
// the persistent grid runs its 148 CTAs in generations, so a single round (one L2 round trip)
// spans the whole current generation plus the tail of the previous one, whose inclusive prefixes
// are already published.
#ifndef VK_LOOK_WIDE
#define VK_LOOK_WIDE 5
#endif
constexpr int kLookWide = VK_LOOK_WIDE;
__device__ __forceinline__ uint32_t look_back(uint64_t* status, uint32_t tile, uint32_t aggregate, uint32_t initial) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {  // `initial`: prefix carried in from outside (sharded scan: the totals of the lower ranks)
    if (lane == 0) status_store(status, ((uint64_t)ST_INCLUSIVE << 32) | (initial + aggregate));
    return initial;
  }
  if (lane == 0) status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_AGGREGATE << 32) | aggregate);
  uint32_t exclusive = 0;
  int top = (int)tile - 1;  // nearest predecessor examined by lane 0
  for (;;) {
    uint64_t s[kLookWide];
#pragma unroll
    for (int i = 0; i < kLookWide; ++i) {
      const int idx = top - 32 * i - lane;
      s[i] = idx >= 0 ? status_load(status + (size_t)idx * kStatusStride) : ((uint64_t)ST_INCLUSIVE << 32);  // before tile 0: prefix 0
    }
    bool done = false;
#pragma unroll
    for (int i = 0; i < kLookWide; ++i) {
      if (done) continue;
      const int idx = top - 32 * i - lane;
      uint32_t spins = 0;
      while ((uint32_t)(s[i] >> 32) == ST_INVALID) {  // predecessor not published yet
        __nanosleep(40);
        s[i] = status_load(status + (size_t)idx * kStatusStride);
        if (++spins > (1u << 25)) __trap();  // seconds without progress: fail loudly instead of hanging the GPU
      }
      const unsigned incl = __ballot_sync(0xFFFFFFFFu, (uint32_t)(s[i] >> 32) == ST_INCLUSIVE);
      if (incl) {
        const int first = __ffs(incl) - 1;  // nearest tile that already knows its inclusive prefix
        exclusive += warp_sum(lane <= first ? (uint32_t)s[i] : 0u);
        done = true;
      } else {
        exclusive += warp_sum((uint32_t)s[i]);
      }
    }
    if (done) break;
    top -= 32 * kLookWide;
  }
  if (lane == 0) status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_INCLUSIVE << 32) | (exclusive + aggregate));
  return exclusive;
}

// ---- lagged look-back (scan_kernel_lag): a tile's prefix is resolved one iteration after its aggregate was published
// First round of status words of `tile`'s look-back window, fetched EARLY and without registers: each lane copies its
// kLookWide status slots (16 bytes each, L2 only: cp.async.cg) into its own row of a shared-memory window, so the
// L2 round trip (1.3-1.5 us while the memory system is saturated) overlaps the local scan of the next tile.
__device__ __forceinline__ void prefetch_window(const uint64_t* status, uint32_t tile, uint64_t* window /* [kLookWide*32][2] */) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kLookWide; ++i) {
    const int idx = (int)tile - 1 - 32 * i - lane;
    uint64_t* dst = window + (size_t)(i * 32 + lane) * 2;
    if (idx >= 0)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst)), "l"(status + (size_t)idx * kStatusStride) : "memory");
    else
      dst[0] = (uint64_t)ST_INCLUSIVE << 32;  // before tile 0: prefix 0
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void load_window(const uint64_t* window, uint64_t (&s)[kLookWide]) {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int i = 0; i < kLookWide; ++i) s[i] = window[(size_t)(i * 32 + lane) * 2];
}
// The look-back walk for a tile whose aggregate was published earlier; `first_round` = the prefetched window.
// dbg (tracing builds only): [0] += polls of a status word that was still INVALID, [1] += look-back rounds
__device__ __forceinline__ uint32_t resolve_prefix(uint64_t* status, uint32_t tile, uint32_t aggregate, uint32_t initial,
                                                   uint64_t (&first_round)[kLookWide], uint32_t* dbg = nullptr) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) return initial;  // tile 0 published its inclusive prefix at once
  uint32_t exclusive = 0;
  int top = (int)tile - 1;
  bool preloaded = true;
  for (;;) {
    if (dbg) dbg[1] += 1u;
    uint64_t s[kLookWide];
#pragma unroll
    for (int i = 0; i < kLookWide; ++i) {
      const int idx = top - 32 * i - lane;
      if (preloaded) s[i] = first_round[i];
      else s[i] = idx >= 0 ? status_load(status + (size_t)idx * kStatusStride) : ((uint64_t)ST_INCLUSIVE << 32);
    }
    preloaded = false;
    bool done = false;
#pragma unroll
    for (int i = 0; i < kLookWide; ++i) {
      if (done) continue;
      const int idx = top - 32 * i - lane;
      uint32_t spins = 0;
      while ((uint32_t)(s[i] >> 32) == ST_INVALID) {
        __nanosleep(40);
        s[i] = status_load(status + (size_t)idx * kStatusStride);
        if (dbg) dbg[0] += 1u;
        if (++spins > (1u << 25)) __trap();
      }
      const unsigned incl = __ballot_sync(0xFFFFFFFFu, (uint32_t)(s[i] >> 32) == ST_INCLUSIVE);
      if (incl) {
        const int first = __ffs(incl) - 1;
        exclusive += warp_sum(lane <= first ? (uint32_t)s[i] : 0u);
        done = true;
      } else {
        exclusive += warp_sum((uint32_t)s[i]);
      }
    }
    if (done) break;
    top -= 32 * kLookWide;
  }
  if (lane == 0) status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_INCLUSIVE << 32) | (exclusive + aggregate));
  return exclusive;
}

// Look-back with an ANCHOR: in a persistent grid of `stride` CTAs the tile `stride` places back is the caller's own
// previous tile, whose inclusive prefix (`own_incl`) it has in a register.  So the walk never needs to FIND an inclusive
// status: it sums the aggregates of the stride - 1 tiles in between (stopping early at an inclusive one if it meets one)
// and adds the anchor.  With 32 * kLookWide >= stride - 1 one window always suffices — no second round, whatever the
// other CTAs have resolved so far.  `first_round`: the prefetched window (distance 32 i + lane + 1 in word i).
__device__ __forceinline__ uint32_t resolve_anchored(uint64_t* status, uint32_t tile, uint32_t aggregate, uint32_t initial,
                                                     uint64_t (&first_round)[kLookWide], uint32_t stride, uint32_t own_incl,
                                                     bool have_own) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) return initial;  // tile 0 published its inclusive prefix at once
  // pass 1: make sure every word that counts is there (not expected to spin: the window is requested iterations after
  // the publish), and find the NEAREST inclusive status, if any.  The ballots are independent of each other.
  uint32_t val[kLookWide];
  uint32_t near_dist = 0xFFFFFFFFu;  // distance of the nearest inclusive predecessor
#pragma unroll
  for (int i = 0; i < kLookWide; ++i) {
    const uint32_t dist = 32u * i + (uint32_t)lane + 1u;
    const bool counts = !have_own || dist < stride;   // tiles at distance >= stride are covered by the anchor
    uint64_t s = first_round[i];
    if (counts && (uint32_t)(s >> 32) == ST_INVALID) {
      const int idx = (int)tile - (int)dist;
      uint32_t spins = 0;
      do {
        __nanosleep(40);
        s = status_load(status + (size_t)idx * kStatusStride);
        if (++spins > (1u << 25)) __trap();
      } while ((uint32_t)(s >> 32) == ST_INVALID);
    }
    val[i] = counts ? (uint32_t)s : 0u;
    const unsigned incl = __ballot_sync(0xFFFFFFFFu, counts && (uint32_t)(s >> 32) == ST_INCLUSIVE);
    if (incl && near_dist == 0xFFFFFFFFu) near_dist = 32u * i + (uint32_t)__ffs(incl);
  }
  // pass 2: every lane adds up its own words up to that distance, ONE warp reduction
  uint32_t mine = 0u;
#pragma unroll
  for (int i = 0; i < kLookWide; ++i) {
    const uint32_t dist = 32u * i + (uint32_t)lane + 1u;
    mine += dist <= near_dist ? val[i] : 0u;
  }
  uint32_t exclusive = warp_sum(mine);
  if (near_dist == 0xFFFFFFFFu) {
    if (!have_own || 32u * (uint32_t)kLookWide + 1u < stride) __trap();  // caller's contract: the window reaches the anchor
    exclusive += own_incl;
  }
  if (lane == 0) status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_INCLUSIVE << 32) | (exclusive + aggregate));
  return exclusive;
}

// ---- TMA bulk-copy / mbarrier helpers (sm_90+; SASS: UBLKCP + SYNCS) -----------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy performed by the TMA unit; completion is signalled on `bar`
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_addr(bar))
               : "memory");
}
// shared -> global bulk copy by the TMA unit (bulk async-group completion); the stores leave the SM without
// occupying LSU store slots, so the issuing warps are not back-pressured while the memory system drains them
__device__ __forceinline__ void tma_store_1d(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_addr(src_smem)), "r"(bytes)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async proxy (before a bulk store reads them)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// plain arrival (release.cta): what the arriving thread wrote to shared memory before is visible to whoever the
// completed phase wakes up
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(done)
                 : "r"(smem_addr(bar)), "r"(parity)
                 : "memory");
    if (!done && ++spins > (1u << 26)) __trap();  // a bulk copy that never lands: fail loudly
  } while (!done);
}
