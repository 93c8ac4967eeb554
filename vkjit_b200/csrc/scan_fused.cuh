// scan_fused.cuh — fused trace -> prefix-sum / compress kernel (NVRTC only; the Makefile embeds this text).
//
// Same single-pass decoupled look-back scan as scan.cu's scan_kernel (persistent CTAs walking tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...; all CTAs co-resident), but the scanned words are not loaded: they are
// computed lane by lane by the generated trace body.  The arrays the trace STREAMS (one word per lane) are
// staged through the same 2-stage TMA ring, one ring slot per streamed array, so a mask such as `x > t`
// costs 4 B/lane of HBM reads and the mask itself never exists in memory (SURVEY.md §8d C28 "fused-mask
// variant": 4 B read + 4p B written instead of 8 + 4p, and no kernel that materialises the mask first).
//
// The generated prelude provides:
//   VK_SCAN_MODE  0 exclusive sum, 1 inclusive sum, 2 compress -> lane indices, 3 compress -> values (root 1)
//   VK_NS         number of streamed arrays (0..6);  VK_VPT  128-bit vectors per thread and tile (6 / max(NS,1))
//   VK_LAG        1: lagged look-back (traces streaming at most one array; see scan.cu: scan_kernel_lag), VK_SLOTS ring slots
//   struct VkPtrs { const u32* s[max(NS,1)]; <gather/scatter pointers> };
//   vk_eval(P, gi, li, in[max(NS,1)], o0, o1)   root words of global lane gi / local lane li
template <bool B> struct VkBool { static constexpr bool value = B; };

// $VKJIT_FSCAN_TRACE=<file> (VK_TRACE): %globaltimer stamps per tile in the spare words 1..11 of the tile's 128-byte
// status line (word 0 is the status): [1] iteration start, [2] tile data landed, [3] evaluated + scanned locally,
// [4] past barrier 1, [5] aggregate published, [6] prefix resolved, [7] a worker warp past barrier 2, [8] its output
// written, [9] past the slot-release barrier, [10] blockIdx, [11] that worker warp reaches barrier 1.
#ifndef VK_TRACE
#define VK_TRACE 0
#endif
#ifndef VK_EARLY
#define VK_EARLY 0
#endif
#ifndef VK_WREG
#define VK_WREG 0
#endif
#ifndef VK_PARK
#define VK_PARK 0
#endif
#ifndef VK_LAGPACK
#define VK_LAGPACK 0
#endif
#if VK_TRACE
// "memory": the timer read must not move across a barrier or the code it brackets
__device__ __forceinline__ unsigned long long vk_stamp_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)::"memory");
  return t;
}
#define VK_STAMP(cond, tile_, w) do { if (cond) status[(size_t)(tile_) * kStatusStride + (w)] = vk_stamp_ns(); } while (0)
#else
#define VK_STAMP(cond, tile_, w) do { } while (0)
#endif
#ifndef VK_CTRL
#define VK_CTRL 0
#endif
#ifndef VK_DIAG
#define VK_DIAG 0
#endif
#ifndef VK_COALESCE
#define VK_COALESCE 0
#endif
#ifndef VK_PACKED
#define VK_PACKED 1
#endif

#if VK_CTRL

// ---- control-warp variant (compress modes, VK_NS <= 1; opt-in: $VKJIT_SCAN_CTRL=1) -----------------------------------------
// EXPERIMENT, measured slower than the lagged kernel below and therefore not the default (profiles/r02_fused_scan.md,
// experiments 6-8); kept because it is parity-tested and its per-phase trace is what explains where the time goes.
// The per-tile trace of the lagged kernel shows a 3.5-3.8 us tile period of which the workers are busy 1.2 us (evaluate
// 0.45, output 0.7): the rest they spend at the two barriers around warp 0's serial section — scan of the tile totals,
// publish, and a look-back that takes 1.2-2 us.  Here a dedicated CONTROL warp (warp VK_T / 32) owns that chain, a
// PRODUCER warp (the next one) issues the TMA refills, and the VK_T worker threads only meet them at mbarriers:
//   worker iteration k:   evaluate tile k -> per-row counts into s_tot[k % R], arrive tot_ready[k % R];
//                         wait resolved[(k - D) % R] -> write tile k - D (flags / row offsets waited in registers) -> free its slot
//   control iteration j:  request the status window of tile j - L + 1 (used in the NEXT iteration); wait tot_ready[j % R]
//                         -> scan the totals of tile j -> publish its aggregate; resolve tile j - L from the window
//                         requested one iteration ago -> s_excl, arrive resolved
// The look-back is ANCHORED (scan_common.cuh: resolve_anchored): the tile `stride` places back is this CTA's own previous
// tile, whose inclusive prefix the control warp has in a register, so one window of stride - 1 aggregates always
// suffices — no search for an inclusive status, no second round.  (VK_COALESCE: a warp row's selected words can leave
// through a 512-byte staging row as 128-byte-aligned coalesced stores; measured slower, off by default.)
// No __syncthreads in the loop: all hand-overs are mbarriers (phase parity = use count of the buffer).  R = 2 D + 1
// buffers: the slowest worker warp can still be writing tile k - 2 D + 1 while the fastest evaluates tile k + 1.
extern "C" __global__ void __launch_bounds__(VK_T + 64, VK_CTAS)
vkjit_trace(const u32 n, const u32 base, const VkPtrs P, u32* __restrict__ out, u32* __restrict__ count_out,
            const u32 num_tiles, uint64_t* __restrict__ state, const u32* __restrict__ initial_ptr,
            const u32* __restrict__ index_base_ptr) {
  constexpr int T = VK_T, VPT = VK_VPT, TILE = T * 4 * VPT, NS = VK_NS, NSA = 1, S = VK_SLOTS, D = VK_DEPTH, L = VK_CLAG, R = 2 * D + 1;
  constexpr int WARPS = T / 32, NTOT = VPT * WARPS, PER_LANE = NTOT / 32;
  constexpr u32 TILE_BYTES = TILE * 4;
  constexpr bool VALUES = VK_SCAN_MODE == 3;
  static_assert(VK_SCAN_MODE >= 2 && NS <= 1 && NTOT % 32 == 0 && VPT <= 4 && L >= 1 && D >= L && (NS == 0 || S >= (VALUES ? D + 2 : 2)), "control-warp fused compress geometry");
  extern __shared__ __align__(128) unsigned char ring_raw[];
  u32* ring = reinterpret_cast<u32*>(ring_raw);          // NS * S input slots
  __shared__ __align__(8) uint64_t full[S], freeb[S], tot_ready[R], resolved[R];
  __shared__ u32 s_tot[R][NTOT];
  __shared__ u32 s_excl[R];
  __shared__ __align__(16) uint64_t s_window[2][kLookWide * 32 * 2];
#if VK_COALESCE
  __shared__ u32 s_stage[WARPS][128];  // one warp row of selected words
#endif

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 first = blockIdx.x, stride = gridDim.x;
  const u32 my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % TILE) != 0;
  const u32 index_base = (VK_SCAN_MODE == 2 && index_base_ptr) ? __ldcg(index_base_ptr) : 0u;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&freeb[s], WARPS); }
#pragma unroll
    for (int r = 0; r < R; ++r) { mbar_init(&tot_ready[r], WARPS); mbar_init(&resolved[r], 1); }
    mbar_fence_init();
  }
  __syncthreads();

  if (warp == WARPS + 1) {
    // ======================================================================================== producer warp (TMA)
    // slot k % S is refilled with tile k + S as soon as every worker warp has released tile k: after its evaluation
    // (the mask words are consumed) or, when the values are re-read for the output, after its output
    if (NS > 0 && lane == 0) {
      auto fill = [&](u32 k) {
        if (k >= my_tiles) return;
        const u32 t = first + k * stride;
        if (ragged && t == num_tiles - 1) return;
        mbar_expect_tx(&full[k % S], TILE_BYTES);
        tma_load_1d(ring + (size_t)(k % S) * TILE, P.s[0] + (size_t)t * TILE, TILE_BYTES, &full[k % S]);
      };
      for (u32 k = 0; k < (u32)S; ++k) fill(k);
      for (u32 q = 0; q + S < my_tiles; ++q) {
        mbar_wait(&freeb[q % S], (q / S) & 1);
        fill(q + S);
      }
    }
    return;
  }
  if (warp == WARPS) {
    // ======================================================================================== control warp
    u32 own_incl = 0u;                                            // inclusive prefix of the tile resolved last
    const bool anchored = 32u * (u32)kLookWide + 1u >= stride;    // one window spans the stride - 1 tiles back to it
    // VK_PACKED: besides its padded status line every tile publishes {AGGREGATE | value} into a PACKED array (8 bytes per
    // tile, behind the status lines).  With the anchor, a tile's prefix is own_incl + the stride - 1 aggregates between the
    // CTA's previous tile and this one — CONTIGUOUS in the packed array: 5 coalesced 16-byte cp.async per lane (20 L2
    // lines) instead of 320 scattered status lines, no search for an inclusive status, one warp reduction.  (Padded lines
    // were introduced against L2-slice hammering by POLLING warps; this window is read once, iterations after the publish.)
    uint64_t* packed = status + (size_t)(num_tiles + 1) * kStatusStride;
    const bool use_packed = VK_PACKED && stride - 1u <= 32u * 2u * (u32)kLookWide - 2u;
    u32 agg_q[L];  // aggregates of tiles j-1 .. j-L
#pragma unroll
    for (int l = 0; l < L; ++l) agg_q[l] = 0u;
    for (u32 j = 0; j < my_tiles + L; ++j) {
      const bool have_cur = j < my_tiles;
      const u32 tj = first + j * stride;  // (trace) stamps of control iteration j go to tile j's status line
      VK_STAMP(have_cur && lane == 0, tj, 1);
      // (1) the status window of tile j - L + 1 is requested ONE iteration before it is used: its predecessors published
      // L - 1 iterations ago, and the round trip (0.3-1.5 us under load) overlaps this iteration's wait for the workers
      {
        const int pj = (int)j - (L - 1);
        if (pj > 0 && (u32)pj < my_tiles && use_packed) {
          // entries [t - stride + 1, t - 1] of the packed array, from an even (16-byte aligned) start
          const u32 t = first + (u32)pj * stride, a0 = (t - stride + 1u) & ~1u;
          uint64_t* dst = s_window[pj & 1];
#pragma unroll
          for (int i = 0; i < kLookWide; ++i) {
            const u32 c = 32u * i + (u32)lane;  // 16-byte chunk: entries a0 + 2c, a0 + 2c + 1
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(dst + 2 * c)), "l"(packed + a0 + 2 * c) : "memory");
          }
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        else if (pj >= 0 && (u32)pj < my_tiles) prefetch_window(status, first + (u32)pj * stride, s_window[pj & 1]);
        else asm volatile("cp.async.commit_group;" ::: "memory");  // keep one group per iteration
      }
      // (2) totals of tile j -> row offsets, aggregate published
      u32 agg_cur = 0u;
      if (have_cur) {
        const u32 tile = first + j * stride;
        mbar_wait(&tot_ready[j % R], (j / R) & 1);
        VK_STAMP(lane == 0, tj, 2);
        u32* tot = s_tot[j % R];
        u32 t[PER_LANE], run = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { t[i] = tot[lane * PER_LANE + i]; run += t[i]; }
        u32 s = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 u = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += u;
        }
        u32 off = s - run;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { tot[lane * PER_LANE + i] = off; off += t[i]; }
        agg_cur = __shfl_sync(0xFFFFFFFFu, s, 31);
        const u32 initial = (tile == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
        if (lane == 0) {
          if (tile == 0) status_store(status, ((uint64_t)ST_INCLUSIVE << 32) | (initial + agg_cur));
          else status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
          if (use_packed) status_store(packed + tile, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
        }
        VK_STAMP(lane == 0, tj, 3);
      }
      // (3) prefix of tile j - L from the window requested in the previous iteration
      if (j >= (u32)L) {
        const u32 rj = j - L, tres = first + rj * stride;
        asm volatile("cp.async.wait_group 1;" ::: "memory");  // everything but this iteration's request has landed
        VK_STAMP(have_cur && lane == 0, tj, 4);
        const u32 init0 = (tres == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
        u32 excl;
        if (use_packed && rj > 0) {
          const u32 lo = tres - stride + 1u, a0 = lo & ~1u;   // entries [lo, tres - 1] count
          const uint64_t* wsrc = s_window[rj & 1];
          u32 mine = 0u;
#pragma unroll
          for (int i = 0; i < kLookWide; ++i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const u32 e = a0 + 2u * (32u * i + (u32)lane) + (u32)h;
              if (e >= lo && e < tres) {
                uint64_t w = wsrc[2 * (32 * i + lane) + h];
                u32 spins = 0;
                while ((u32)(w >> 32) == ST_INVALID) {  // (not expected: requested iterations after the publish)
                  __nanosleep(40);
                  w = status_load(packed + e);
                  if (++spins > (1u << 25)) __trap();
                }
                mine += (u32)w;
              }
            }
          }
          excl = own_incl + warp_sum(mine);
          if (lane == 0) status_store(status + (size_t)tres * kStatusStride, ((uint64_t)ST_INCLUSIVE << 32) | (excl + agg_q[L - 1]));
        } else {
          uint64_t window[kLookWide];
          const uint64_t* wsrc = s_window[rj & 1];
#pragma unroll
          for (int i = 0; i < kLookWide; ++i) window[i] = wsrc[(size_t)(i * 32 + lane) * 2];
          // the anchor: this CTA's previous tile (tres - stride), whose inclusive prefix is in own_incl
          excl = anchored ? resolve_anchored(status, tres, agg_q[L - 1], init0, window, stride, own_incl, rj > 0)
                          : resolve_prefix(status, tres, agg_q[L - 1], init0, window);
        }
        own_incl = excl + agg_q[L - 1];
        if (lane == 0) {
          s_excl[rj % R] = excl;
          if (tres == num_tiles - 1) *count_out = excl + agg_q[L - 1];
        }
        __syncwarp();  // every lane's row offsets of tile rj (written L iterations ago) and lane 0's prefix are in place
        if (lane == 0) mbar_arrive(&resolved[rj % R]);
        VK_STAMP(have_cur && lane == 0, tj, 5);
      }
#pragma unroll
      for (int l = L - 1; l > 0; --l) agg_q[l] = agg_q[l - 1];
      agg_q[0] = agg_cur;
      VK_STAMP(have_cur && lane == 0, tj, 6);
#if VK_TRACE
      if (have_cur && lane == 0) status[(size_t)tj * kStatusStride + 15] = blockIdx.x;
#endif
      __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return;
  }

  // ============================================================================================ workers
  auto fetch = [&](int j, bool staged, u32 slot, size_t e, u32 (&w)[4]) {
    if (NS == 0) return;
    if (staged) {
      const uint4 a = reinterpret_cast<const uint4*>(ring + (size_t)slot * TILE)[j * T + threadIdx.x];
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = e + c < n ? P.s[0][e + c] : 0u;
    }
  };
  u32 flags_q[D], pre_q[D];  // tiles k-1 .. k-D: 4 selection bits / one 8-bit exclusive row offset per vector
#pragma unroll
  for (int d = 0; d < D; ++d) { flags_q[d] = 0u; pre_q[d] = 0u; }

  for (u32 k = 0; k < my_tiles + D; ++k) {
    const bool have_cur = k < my_tiles, have_out = k >= (u32)D;
    u32 flags_c = 0u, pre_c = 0u;
    if (have_cur) {  // ---- evaluate tile k, count per warp row
      const u32 tile = first + k * stride;
      const size_t tile_base = (size_t)tile * TILE;
      const bool whole = !(ragged && tile == num_tiles - 1);
      VK_STAMP(threadIdx.x == 64, tile, 7);
      if (whole && NS > 0) mbar_wait(&full[k % S], (k / S) & 1);
      VK_STAMP(threadIdx.x == 64, tile, 8);
      u32 own = 0u;
      auto evaluate = [&](auto whole_c) {
        constexpr bool W = decltype(whole_c)::value;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          u32 w[4];
          fetch(j, W, k % S, e, w);
          u32 r[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            u32 o0 = 0u, o1 = 0u;
            if (W || e + c < n) {
              u32 in[NSA];
              in[0] = NS > 0 ? w[c] : 0u;
              vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, o1);
            }
            r[c] = o0;
          }
          const u32 f = r[0] | (r[1] << 1) | (r[2] << 2) | (r[3] << 3);  // the word of a Bool root is exactly 0 or 1
          flags_c |= f << (4 * j);
          own |= (u32)__popc(f) << (8 * j);
        }
      };
      if (whole) evaluate(VkBool<true>{}); else evaluate(VkBool<false>{});
      u32 pk = own;  // four 8-bit per-slot counts, one packed warp scan
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xFFFFFFFFu, pk, o);
        if (lane >= o) pk += t;
      }
      pre_c = pk - own;
      if (lane == 31) {
        u32* tot = s_tot[k % R];
#pragma unroll
        for (int j = 0; j < VPT; ++j) tot[j * WARPS + warp] = (pk >> (8 * j)) & 0xFFu;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&tot_ready[k % R]);
        if (!VALUES && NS > 0) mbar_arrive(&freeb[k % S]);  // the mask words are consumed
      }
      VK_STAMP(threadIdx.x == 64, tile, 9);
    }
    if (have_out) {  // ---- selected lanes of tile k - D, written at their rank
      const u32 ko = k - D, tout = first + ko * stride;
      const size_t tile_base = (size_t)tout * TILE;
      const bool whole = !(ragged && tout == num_tiles - 1);
      VK_STAMP(threadIdx.x == 64, tout, 10);
      mbar_wait(&resolved[ko % R], (ko / R) & 1);
      VK_STAMP(threadIdx.x == 64, tout, 11);
      const u32 tile_excl = s_excl[ko % R];
      const u32* tot = s_tot[ko % R];
      const u32 flags_o = flags_q[D - 1], pre_o = pre_q[D - 1];
      auto emit = [&](auto whole_c) {
        constexpr bool W = decltype(whole_c)::value;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const u32 f = (flags_o >> (4 * j)) & 0xFu;
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          u32 v[4] = {0u, 0u, 0u, 0u};
          if (f) {
            if (VALUES) {  // all four lanes of the vector are re-evaluated (no per-lane branches)
              u32 w[4];
              fetch(j, W, ko % S, e, w);
#pragma unroll
              for (int c = 0; c < 4; ++c) {
                if (W || e + c < n) {
                  u32 in[NSA];
                  in[0] = NS > 0 ? w[c] : 0u;
                  u32 o0;
                  vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, v[c]);
                }
              }
            } else {
#pragma unroll
              for (int c = 0; c < 4; ++c) v[c] = index_base + (u32)(e + c);
            }
          }
          const u32 off = (pre_o >> (8 * j)) & 0xFFu;  // rank of this vector's first selected lane within the warp row
          const u32 start = tile_excl + tot[j * WARPS + warp];
          const u32 s1 = f & 1u, s2 = s1 + ((f >> 1) & 1u), s3 = s2 + ((f >> 2) & 1u);
#if VK_DIAG == 1   /* timing diagnostic (wrong results): no output stores except an impossible one that keeps the values alive */
          if ((v[0] ^ v[1] ^ v[2] ^ v[3]) == 0x9E3779B9u && f == 0x1Fu) out[start + off + s1 + s2 + s3] = v[0];
#elif VK_COALESCE
          // The selected lanes of a warp row (128 consecutive lanes) form ONE contiguous run of the output.  Written lane by
          // lane, every 32-byte sector of it is hit by several predicated store instructions; here the run is packed in a
          // 512-byte per-warp staging row first and leaves as 128-byte-aligned, fully coalesced stores.  Measured SLOWER
          // (profiles/r02_fused_scan.md, experiment 8; ncu: the L2 runs at 29 % with the lane-by-lane stores, it was never
          // the limit) — so it is only compiled in under $VKJIT_FSCAN_DIAG=2; the default writes lane by lane.
          u32* stg = s_stage[warp];
          if (f & 1u) stg[off] = v[0];
          if (f & 2u) stg[off + s1] = v[1];
          if (f & 4u) stg[off + s2] = v[2];
          if (f & 8u) stg[off + s3] = v[3];
          const u32 len = __shfl_sync(0xFFFFFFFFu, off + s3 + ((f >> 3) & 1u), 31);  // selected lanes in the row
          __syncwarp();
          const u32 a = start & 31u;
          for (u32 c = 0; c * 32u < a + len; ++c) {
            const int idx = (int)(c * 32u + (u32)lane) - (int)a;
            if (idx >= 0 && (u32)idx < len) out[start + (u32)idx] = stg[idx];
          }
          __syncwarp();
#else
          u32* q = out + (start + off);
          if (f & 1u) q[0] = v[0];
          if (f & 2u) q[s1] = v[1];
          if (f & 4u) q[s2] = v[2];
          if (f & 8u) q[s3] = v[3];
#endif
        }
      };
      if (whole) emit(VkBool<true>{}); else emit(VkBool<false>{});
      if (VALUES && NS > 0) {  // the slot of tile k - D was kept for the re-evaluation: free now
        __syncwarp();
        if (lane == 0) mbar_arrive(&freeb[ko % S]);
      }
      VK_STAMP(threadIdx.x == 64, tout, 12);
    }
#pragma unroll
    for (int d = D - 1; d > 0; --d) { flags_q[d] = flags_q[d - 1]; pre_q[d] = pre_q[d - 1]; }
    flags_q[0] = flags_c; pre_q[0] = pre_c;
  }
}

#elif !VK_LAG

extern "C" __global__ void __launch_bounds__(VK_T, 1024 / VK_T)
vkjit_trace(const u32 n, const u32 base, const VkPtrs P, u32* __restrict__ out, u32* __restrict__ count_out,
            const u32 num_tiles, uint64_t* __restrict__ state, const u32* __restrict__ initial_ptr,
            const u32* __restrict__ index_base_ptr) {  // mode 2: global index of lane 0 (sharded masks)
  constexpr int T = VK_T, VPT = VK_VPT, TILE = T * 4 * VPT, NS = VK_NS, NSA = NS > 0 ? NS : 1, S = 2;
  constexpr int WARPS = T / 32, NTOT = VPT * WARPS, PER_LANE = (NTOT + 31) / 32;
  constexpr u32 TILE_BYTES = TILE * 4;
  constexpr bool COMPRESS = VK_SCAN_MODE >= 2, VALUES = VK_SCAN_MODE == 3;
  // compress -> values re-evaluates the selected lanes when it writes them, so the ring slot is kept
  // until the tile's output phase is over (the other slot is still being filled meanwhile)
  constexpr bool LATE_RELEASE = VALUES && NS > 0;
  extern __shared__ __align__(128) unsigned char ring_raw[];
  u32* ring = reinterpret_cast<u32*>(ring_raw);  // [stage][stream][TILE]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ u32 s_tot[2][NTOT];
  __shared__ u32 s_tile_excl[2];

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 first = blockIdx.x, stride = gridDim.x;
  const u32 my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % TILE) != 0;  // the globally last tile is partial: guarded loads, no TMA
  const u32 index_base = (VK_SCAN_MODE == 2 && index_base_ptr) ? __ldcg(index_base_ptr) : 0u;

  auto fill = [&](u32 t, int stage) {  // thread 0: one bulk copy per streamed array, all on the stage's barrier
    if (ragged && t == num_tiles - 1) return;
    mbar_expect_tx(&full[stage], (u32)NS * TILE_BYTES);
#pragma unroll
    for (int s = 0; s < NS; ++s)
      tma_load_1d(ring + ((size_t)stage * NS + s) * TILE, P.s[s] + (size_t)t * TILE, TILE_BYTES, &full[stage]);
  };
  // words of vector j of the current tile, per streamed array
  auto fetch = [&](int j, bool staged, int stage, size_t e, u32 (&w)[NSA][4]) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
      if (staged) {
        const uint4 a = reinterpret_cast<const uint4*>(ring + ((size_t)stage * NS + s) * TILE)[j * T + threadIdx.x];
        w[s][0] = a.x; w[s][1] = a.y; w[s][2] = a.z; w[s][3] = a.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; ++c) w[s][c] = e + c < n ? P.s[s][e + c] : 0u;
      }
    }
  };

  if (NS > 0) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (u32 k = 0; k < (u32)S && k < my_tiles; ++k) fill(first + k * stride, (int)k);
  }

  for (u32 k = 0; k < my_tiles; ++k) {
    const u32 tile = first + k * stride;
    const int stage = k % S;
    const int buf = k & 1;
    const size_t tile_base = (size_t)tile * TILE;
    const bool whole = !(ragged && tile == num_tiles - 1);
    const bool staged = whole && NS > 0;

    if (staged) mbar_wait(&full[stage], (k / S) & 1);
    uint4 x[VPT];     // sums: the addends
    u32 flags[VPT];   // compress: 4 selection bits per vector
    // W: whole tile (no bounds checks; the common case is compiled without them)
    auto evaluate = [&](auto whole_c) {
      constexpr bool W = decltype(whole_c)::value;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
        u32 w[NSA][4];
        fetch(j, W && NS > 0, stage, e, w);
        u32 r[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          u32 o0 = 0u, o1 = 0u;
          if (W || e + c < n) {  // lanes past the end are never evaluated (their gathers would be out of range)
            u32 in[NSA];
#pragma unroll
            for (int s = 0; s < NS; ++s) in[s] = w[s][c];
            vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, o1);
          }
          r[c] = o0;
        }
        // the word of a Bool root is exactly 0 or 1
        if (COMPRESS) flags[j] = r[0] | (r[1] << 1) | (r[2] << 2) | (r[3] << 3);
        else { x[j].x = r[0]; x[j].y = r[1]; x[j].z = r[2]; x[j].w = r[3]; }
      }
    };
    if (whole) evaluate(VkBool<true>{}); else evaluate(VkBool<false>{});

    // 1) per-vector sums, 2) inclusive warp scan per register slot, 3) one warp scans the
    // (slot, warp) totals in tile order, 4) look-back gives the tile's global offset.
    u32 vsum[VPT], wincl[VPT];
    if (COMPRESS) {
      // a warp's inclusive count per slot is at most 32 x 4 = 128: four slots share one 32-bit word (8-bit
      // fields, no carries between them), so VPT warp scans become ceil(VPT / 4)
      constexpr int G = (VPT + 3) / 4;
      u32 pk[G];
#pragma unroll
      for (int g = 0; g < G; ++g) pk[g] = 0u;
#pragma unroll
      for (int j = 0; j < VPT; ++j) { vsum[j] = (u32)__popc(flags[j]); pk[j / 4] |= vsum[j] << (8 * (j % 4)); }
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_up_sync(0xFFFFFFFFu, pk[g], o);
          if (lane >= o) pk[g] += t;
        }
      }
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        wincl[j] = (pk[j / 4] >> (8 * (j % 4))) & 0xFFu;
        if (lane == 31) s_tot[buf][j * WARPS + warp] = wincl[j];
      }
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        vsum[j] = x[j].x + x[j].y + x[j].z + x[j].w;
        u32 s = vsum[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += t;
        }
        wincl[j] = s;
        if (lane == 31) s_tot[buf][j * WARPS + warp] = s;
      }
    }
    __syncthreads();  // every thread has consumed its part of ring[stage]
    if (NS > 0 && !LATE_RELEASE && threadIdx.x == 0 && k + S < my_tiles) fill(first + (k + S) * stride, stage);
    if (warp == 0) {
      u32 t[PER_LANE], run = 0;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) { t[i] = lane * PER_LANE + i < NTOT ? s_tot[buf][lane * PER_LANE + i] : 0u; run += t[i]; }
      u32 s = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const u32 u = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if (lane >= o) s += u;
      }
      u32 off = s - run;  // exclusive offset of this lane's first entry
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) { if (lane * PER_LANE + i < NTOT) s_tot[buf][lane * PER_LANE + i] = off; off += t[i]; }
      const u32 aggregate = __shfl_sync(0xFFFFFFFFu, s, 31);
      const u32 initial = (tile == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
      const u32 excl = look_back(status, tile, aggregate, initial);
      if (lane == 0) {
        s_tile_excl[buf] = excl;
        if (COMPRESS && tile == num_tiles - 1) *count_out = excl + aggregate;
      }
    }
    __syncthreads();
    const u32 tile_excl = s_tile_excl[buf];

    auto emit = [&](auto whole_c) {
      constexpr bool W = decltype(whole_c)::value;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
        const u32 p = tile_excl + s_tot[buf][j * WARPS + warp] + (wincl[j] - vsum[j]);  // exclusive prefix of lane e
        if (!COMPRESS) {
          uint4 r;
          if (VK_SCAN_MODE == 0) { r.x = p; r.y = p + x[j].x; r.z = r.y + x[j].y; r.w = r.z + x[j].z; }
          else { r.x = p + x[j].x; r.y = r.x + x[j].y; r.z = r.y + x[j].z; r.w = r.z + x[j].w; }
          if (W || e + 3 < n) st_stream(reinterpret_cast<uint4*>(out + e), r);
          else {
            if (e + 0 < n) out[e + 0] = r.x;
            if (e + 1 < n) out[e + 1] = r.y;
            if (e + 2 < n) out[e + 2] = r.z;
          }
        } else if (flags[j]) {
          // selected lanes are written at their rank; flags of out-of-range lanes are 0
          const u32 f = flags[j];
          u32 v[4];
          if (VALUES) {
            u32 w[NSA][4];
            fetch(j, W && NS > 0, stage, e, w);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              v[c] = 0u;
              if (W || e + c < n) {  // all four lanes of the vector are re-evaluated (no per-lane branches)
                u32 in[NSA];
#pragma unroll
                for (int s = 0; s < NS; ++s) in[s] = w[s][c];
                u32 o0;
                vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, v[c]);
              }
            }
          } else {
#pragma unroll
            for (int c = 0; c < 4; ++c) v[c] = index_base + (u32)(e + c);
          }
          u32* q = out + p;  // one 64-bit address per vector; the slots of its lanes follow from the flag bits
          const u32 s1 = f & 1u, s2 = s1 + ((f >> 1) & 1u), s3 = s2 + ((f >> 2) & 1u);
          if (f & 1u) q[0] = v[0];
          if (f & 2u) q[s1] = v[1];
          if (f & 4u) q[s2] = v[2];
          if (f & 8u) q[s3] = v[3];
        }
      }
    };
    if (whole) emit(VkBool<true>{}); else emit(VkBool<false>{});
    if (LATE_RELEASE) {
      __syncthreads();  // the output phase read ring[stage] again
      if (threadIdx.x == 0 && k + S < my_tiles) fill(first + (k + S) * stride, stage);
    }
    // s_tot/s_tile_excl are double-buffered: iteration k+2 rewrites buffer `buf` only after every
    // thread passed the first barrier of iteration k+1, i.e. after it finished reading it here.
  }
}

#else  // VK_LAG

// ---- lagged variant: tile k's aggregate is published right after its evaluation + local scan; its prefix is resolved and
// its output written one iteration later (status window prefetched with cp.async before the next tile is evaluated).
// VK_NS <= 1.  Modes 0/1: row-relative results wait in registers and leave through a staging tile + one TMA bulk store.
// Mode 2: 4 selection bits + an 8-bit row offset per vector wait in two registers.  Mode 3: additionally the tile's ring slot
// stays resident one more iteration (VK_SLOTS = 4) and the selected lanes' values are re-evaluated from it.
extern "C" __global__ void __launch_bounds__(VK_T, 1024 / VK_T)
vkjit_trace(const u32 n, const u32 base, const VkPtrs P, u32* __restrict__ out, u32* __restrict__ count_out,
            const u32 num_tiles, uint64_t* __restrict__ state, const u32* __restrict__ initial_ptr,
            const u32* __restrict__ index_base_ptr) {
  constexpr int T = VK_T, VPT = VK_VPT, TILE = T * 4 * VPT, NS = VK_NS, NSA = 1, S = VK_SLOTS;
  constexpr int TW = 64;  // the traced worker thread (warp 2)
  constexpr int WARPS = T / 32, NTOT = VPT * WARPS, PER_LANE = NTOT / 32;
  constexpr u32 TILE_BYTES = TILE * 4;
  constexpr bool COMPRESS = VK_SCAN_MODE >= 2, VALUES = VK_SCAN_MODE == 3;
  static_assert(NS <= 1 && NTOT % 32 == 0 && (!COMPRESS || VPT <= 4), "lagged fused scan geometry");
  extern __shared__ __align__(128) unsigned char ring_raw[];
  u32* ring = reinterpret_cast<u32*>(ring_raw);          // NS * S input slots
  u32* stage_out = ring + (size_t)NS * S * TILE;         // modes 0/1: output staging tile
  // VK_PARK (prefix sums of traces that stream nothing): there is no input ring, so the whole shared memory can hold
  // results instead — every thread PARKS its row-relative results of tile k in one of two tile-sized buffers and adds the
  // tile's prefix to them one iteration later, straight into 128-bit stores.  No registers are held across the
  // look-back, so the tile can be as large as the immediate-look-back kernel's (VPT 6) while the look-back still lags.
  constexpr bool PARK = VK_PARK != 0;
  static_assert(!PARK || (NS == 0 && !COMPRESS), "parked results: prefix sums without streamed inputs");
  uint4* park = reinterpret_cast<uint4*>(ring_raw);      // [2][TILE / 4]
  __shared__ __align__(8) uint64_t full[S];
  __shared__ u32 s_tot[3][NTOT];
  __shared__ u32 s_tile_excl;
  __shared__ __align__(16) uint64_t s_window[kLookWide * 32 * 2];

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const u32 first = blockIdx.x, stride = gridDim.x;
  const u32 my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % TILE) != 0;
  const u32 index_base = (VK_SCAN_MODE == 2 && index_base_ptr) ? __ldcg(index_base_ptr) : 0u;

  auto fill = [&](u32 k) {  // one thread: the streamed array's tile k into slot k % S
    if (NS == 0 || k >= my_tiles) return;
    const u32 t = first + k * stride;
    if (ragged && t == num_tiles - 1) return;
    mbar_expect_tx(&full[k % S], TILE_BYTES);
    tma_load_1d(ring + (size_t)(k % S) * TILE, P.s[0] + (size_t)t * TILE, TILE_BYTES, &full[k % S]);
  };
  // words of vector j of tile (slot, e) of the streamed array
  auto fetch = [&](int j, bool staged, u32 slot, size_t e, u32 (&w)[4]) {
    if (NS == 0) return;
    if (staged) {
      const uint4 a = reinterpret_cast<const uint4*>(ring + (size_t)slot * TILE)[j * T + threadIdx.x];
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w;
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = e + c < n ? P.s[0][e + c] : 0u;
    }
  };

  if (NS > 0) {
    if (threadIdx.x == 0) {
#pragma unroll
      for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
      mbar_fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0)
      for (u32 k = 0; k < (u32)S; ++k) fill(k);
  }

  uint4 xp[PARK ? 1 : VPT];     // modes 0/1, previous tile: results relative to the start of the vector's warp row
  u32 flags_p = 0u, pre_p = 0u; // modes 2/3, previous tile: selection bits / 8-bit exclusive row offsets
  u32 agg_prev = 0u;            // warp 0: aggregate of the previous tile
  // VK_LAGPACK: anchored look-back over a PACKED copy of the tile aggregates (8 bytes per tile behind the status lines).
  // The tile `stride` places back is this CTA's own previous tile, whose inclusive prefix warp 0 keeps in own_incl; the
  // stride - 1 aggregates in between are contiguous in the packed array: the same 5 cp.async per lane as for the padded
  // status window, but 20 L2 lines instead of 160, a window that always reaches the anchor (296 tiles per generation
  // with two CTAs per SM) — so never a second round and no search for an inclusive status.  The first tile of a CTA
  // takes the generic walk over the padded status lines.
  u32 own_incl = 0u;
  uint64_t* packed = status + (size_t)(num_tiles + 1) * kStatusStride;
  const bool use_packed = VK_LAGPACK && !VK_WREG && !VK_EARLY && stride - 1u <= 32u * 2u * (u32)kLookWide - 2u;
#pragma unroll
  for (int j = 0; j < (PARK ? 1 : VPT); ++j) xp[j] = make_uint4(0u, 0u, 0u, 0u);

  for (u32 k = 0; k <= my_tiles; ++k) {
    const bool have_cur = k < my_tiles, have_prev = k > 0;
    const u32 tile = first + k * stride;
    uint4 xc[VPT];
    u32 flags_c = 0u, pre_c = 0u;
#pragma unroll
    for (int j = 0; j < VPT; ++j) xc[j] = make_uint4(0u, 0u, 0u, 0u);
    // status window of the tile that is resolved in THIS iteration (tile k-1).  VK_EARLY (experiment, measured slower):
    // requested one phase earlier, right after the previous iteration's resolve (below).
    // VK_WREG: the window is read with STRONG loads (ld.relaxed.gpu) into registers instead of cp.async.cg into shared
    // memory: a weak load may be served a stale copy of a status line, and the per-tile trace showed a third of all
    // tiles polling words again that had been published microseconds earlier (profiles/r02_fused_scan.md, experiment 5)
    uint64_t wreg[kLookWide];
    if (VK_WREG && warp == 0 && have_prev) {
#pragma unroll
      for (int i = 0; i < kLookWide; ++i) {
        const int idx = (int)(tile - stride) - 1 - 32 * i - lane;
        wreg[i] = idx >= 0 ? status_load(status + (size_t)idx * kStatusStride) : ((uint64_t)ST_INCLUSIVE << 32);
      }
    }
    if (use_packed && warp == 0 && k > 1) {   // tile k-1 has a predecessor in this CTA: the packed, anchored window
      const u32 t = tile - stride, a0 = (t - stride + 1u) & ~1u;   // entries [t - stride + 1, t - 1], from an even start
#pragma unroll
      for (int i = 0; i < kLookWide; ++i) {
        const u32 c = 32u * i + (u32)lane;  // 16-byte chunk: entries a0 + 2c, a0 + 2c + 1
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(s_window + 2 * c)), "l"(packed + a0 + 2 * c) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    } else
    if (!VK_WREG && !VK_EARLY && warp == 0 && have_prev) prefetch_window(status, tile - stride, s_window);
    if (have_cur) {  // ---- evaluate tile k and scan it locally
      const size_t tile_base = (size_t)tile * TILE;
      const bool whole = !(ragged && tile == num_tiles - 1);
      VK_STAMP(threadIdx.x == 0, tile, 1);
      if (whole && NS > 0) mbar_wait(&full[k % S], (k / S) & 1);
      VK_STAMP(threadIdx.x == 0, tile, 2);
      u32 own = 0u;
      auto evaluate = [&](auto whole_c) {
        constexpr bool W = decltype(whole_c)::value;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          u32 w[4];
          fetch(j, W, k % S, e, w);
          u32 r[4];
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            u32 o0 = 0u, o1 = 0u;
            if (W || e + c < n) {
              u32 in[NSA];
              in[0] = NS > 0 ? w[c] : 0u;
              vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, o1);
            }
            r[c] = o0;
          }
          if (COMPRESS) {
            const u32 f = r[0] | (r[1] << 1) | (r[2] << 2) | (r[3] << 3);  // the word of a Bool root is exactly 0 or 1
            flags_c |= f << (4 * j);
            own |= (u32)__popc(f) << (8 * j);
          } else {
            xc[j].x = r[0]; xc[j].y = r[1]; xc[j].z = r[2]; xc[j].w = r[3];
          }
        }
      };
      if (whole) evaluate(VkBool<true>{}); else evaluate(VkBool<false>{});
      u32* tot = s_tot[k % 3];
      if (COMPRESS) {
        u32 pk = own;  // four 8-bit per-slot counts, one packed warp scan
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 t = __shfl_up_sync(0xFFFFFFFFu, pk, o);
          if (lane >= o) pk += t;
        }
        pre_c = pk - own;
        if (lane == 31) {
#pragma unroll
          for (int j = 0; j < VPT; ++j) tot[j * WARPS + warp] = (pk >> (8 * j)) & 0xFFu;
        }
      } else {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const uint4 a = xc[j];
          const u32 vs = a.x + a.y + a.z + a.w;
          u32 s = vs;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const u32 t = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (lane >= o) s += t;
          }
          if (lane == 31) tot[j * WARPS + warp] = s;
          const u32 p = s - vs;
          if (VK_SCAN_MODE == 0) { xc[j].x = p; xc[j].y = p + a.x; xc[j].z = xc[j].y + a.y; xc[j].w = xc[j].z + a.z; }
          else { xc[j].x = p + a.x; xc[j].y = xc[j].x + a.y; xc[j].z = xc[j].y + a.z; xc[j].w = xc[j].z + a.w; }
          if (PARK) park[(size_t)(k & 1) * (TILE / 4) + j * T + threadIdx.x] = xc[j];  // read back by this same thread
        }
      }
    }
    VK_STAMP(have_cur && threadIdx.x == 0, tile, 3);
    VK_STAMP(have_cur && threadIdx.x == TW, tile, 11);
    __syncthreads();  // s_tot[k % 3] complete; (modes 0-2) every thread has consumed ring slot k % S
    VK_STAMP(have_cur && threadIdx.x == 0, tile, 4);
#if VK_TRACE
    if (have_cur && threadIdx.x == 0) status[(size_t)tile * kStatusStride + 10] = blockIdx.x;
#endif
    if (!VALUES && have_cur && threadIdx.x == 32) fill(k + S);
    if (warp == 0) {
      u32 agg_cur = 0u;
      if (have_cur) {
        u32* tot = s_tot[k % 3];
        u32 t[PER_LANE], run = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { t[i] = tot[lane * PER_LANE + i]; run += t[i]; }
        u32 s = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const u32 u = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += u;
        }
        u32 off = s - run;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { tot[lane * PER_LANE + i] = off; off += t[i]; }
        agg_cur = __shfl_sync(0xFFFFFFFFu, s, 31);
        const u32 initial = (tile == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
        if (lane == 0) {
          if (tile == 0) status_store(status, ((uint64_t)ST_INCLUSIVE << 32) | (initial + agg_cur));
          else status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
          if (use_packed) status_store(packed + tile, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
        }
        VK_STAMP(lane == 0, tile, 5);
      }
      if (have_prev && use_packed && k > 1) {
        const u32 tprev = tile - stride, lo = tprev - stride + 1u, a0 = lo & ~1u;
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        VK_STAMP(lane == 0, tprev, 12);
        u32 mine = 0u, polled = 0u;
#pragma unroll
        for (int i = 0; i < kLookWide; ++i) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const u32 e = a0 + 2u * (32u * i + (u32)lane) + (u32)h;
            if (e >= lo && e < tprev) {
              uint64_t w = s_window[2 * (32 * i + lane) + h];
              u32 spins = 0;
              while ((u32)(w >> 32) == ST_INVALID) {  // a predecessor that runs more than a tile period behind
                __nanosleep(40);
                w = status_load(packed + e);
                polled += 1u;
                if (++spins > (1u << 25)) __trap();
              }
              mine += (u32)w;
            }
          }
        }
        const u32 excl = own_incl + warp_sum(mine);
        own_incl = excl + agg_prev;
        if (lane == 0) {
          status_store(status + (size_t)tprev * kStatusStride, ((uint64_t)ST_INCLUSIVE << 32) | own_incl);
          s_tile_excl = excl;
          if (COMPRESS && tprev == num_tiles - 1) *count_out = excl + agg_prev;
        }
#if VK_TRACE
        { const u32 polls = __reduce_add_sync(0xFFFFFFFFu, polled);
          if (lane == 0) { status[(size_t)tprev * kStatusStride + 13] = polls; status[(size_t)tprev * kStatusStride + 14] = 1u; } }
#endif
        VK_STAMP(lane == 0, tprev, 6);
      } else
      if (have_prev) {
        const u32 tprev = tile - stride;
        uint64_t window[kLookWide];
        if (VK_WREG) {
#pragma unroll
          for (int i = 0; i < kLookWide; ++i) window[i] = wreg[i];
        } else {
          load_window(s_window, window);
        }
        VK_STAMP(lane == 0, tprev, 12);   // the prefetched window has landed
#if VK_TRACE
        u32 dbg[2] = {0u, 0u};
        const u32 excl = resolve_prefix(status, tprev, agg_prev, (tprev == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u, window, dbg);
        // [13] INVALID polls summed over the warp's lanes, [14] look-back rounds
        { const u32 polls = __reduce_add_sync(0xFFFFFFFFu, dbg[0]);
          if (lane == 0) { status[(size_t)tprev * kStatusStride + 13] = polls; status[(size_t)tprev * kStatusStride + 14] = dbg[1]; } }
#else
        const u32 excl = resolve_prefix(status, tprev, agg_prev, (tprev == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u, window);
#endif
        own_incl = excl + agg_prev;
        if (lane == 0) {
          s_tile_excl = excl;
          if (COMPRESS && tprev == num_tiles - 1) *count_out = excl + agg_prev;
        }
        VK_STAMP(lane == 0, tprev, 6);
      }
      agg_prev = agg_cur;
    }
    if (!COMPRESS && !PARK && threadIdx.x == 0) tma_store_wait_read();  // the previous bulk store has read the staging tile
    __syncthreads();
    // tile k's aggregate was published a moment ago; its window is read now, ~1 us before it is needed, while the
    // predecessors' aggregates of the same generation (published at about the same time as ours) become visible
    if (!VK_WREG && VK_EARLY && warp == 0 && have_cur) prefetch_window(status, tile, s_window);
    VK_STAMP(have_prev && threadIdx.x == TW, tile - stride, 7);
    if (have_prev) {  // ---- output of tile k-1
      const u32 kp = k - 1, tprev = tile - stride;
      const size_t tile_base = (size_t)tprev * TILE;
      const bool whole = !(ragged && tprev == num_tiles - 1);
      const u32 tile_excl = s_tile_excl;
      const u32* tot = s_tot[kp % 3];
      if (!COMPRESS) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          const u32 p = tile_excl + tot[j * WARPS + warp];
          uint4 r;
          const uint4 x = PARK ? park[(size_t)(kp & 1) * (TILE / 4) + j * T + threadIdx.x] : xp[PARK ? 0 : j];
          r.x = x.x + p; r.y = x.y + p; r.z = x.z + p; r.w = x.w + p;
          if (whole && !PARK) reinterpret_cast<uint4*>(stage_out)[j * T + threadIdx.x] = r;
          else if (e + 3 < n) st_stream(reinterpret_cast<uint4*>(out + e), r);
          else {
            if (e + 0 < n) out[e + 0] = r.x;
            if (e + 1 < n) out[e + 1] = r.y;
            if (e + 2 < n) out[e + 2] = r.z;
          }
        }
        if (whole && !PARK) {
          fence_proxy_async();
          __syncthreads();
          if (threadIdx.x == 0) tma_store_1d(out + tile_base, stage_out, TILE_BYTES);
        }
      } else {
        auto emit = [&](auto whole_c) {
          constexpr bool W = decltype(whole_c)::value;
#pragma unroll
          for (int j = 0; j < VPT; ++j) {
            const u32 f = (flags_p >> (4 * j)) & 0xFu;
            if (f) {
              const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
              u32 v[4];
              if (VALUES) {  // all four lanes of the vector are re-evaluated (no per-lane branches)
                u32 w[4];
                fetch(j, W, kp % S, e, w);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  v[c] = 0u;
                  if (W || e + c < n) {
                    u32 in[NSA];
                    in[0] = NS > 0 ? w[c] : 0u;
                    u32 o0;
                    vk_eval(P, base + (u32)(e + c), (u32)(e + c), in, o0, v[c]);
                  }
                }
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) v[c] = index_base + (u32)(e + c);
              }
              u32* q = out + (tile_excl + tot[j * WARPS + warp] + ((pre_p >> (8 * j)) & 0xFFu));
              const u32 s1 = f & 1u, s2 = s1 + ((f >> 1) & 1u), s3 = s2 + ((f >> 2) & 1u);
              if (f & 1u) q[0] = v[0];
              if (f & 2u) q[s1] = v[1];
              if (f & 4u) q[s2] = v[2];
              if (f & 8u) q[s3] = v[3];
            }
          }
        };
        if (whole) emit(VkBool<true>{}); else emit(VkBool<false>{});
        VK_STAMP(threadIdx.x == TW, tprev, 8);
        if (VALUES && NS > 0) {  // the slot of tile k-1 was kept for the re-evaluation: free now
          __syncthreads();
          if (threadIdx.x == 32) fill(kp + S);
        }
        VK_STAMP(threadIdx.x == TW, tprev, 9);
      }
    }
    if (!PARK) {
#pragma unroll
      for (int j = 0; j < VPT; ++j) xp[PARK ? 0 : j] = xc[j];
    }
    flags_p = flags_c; pre_p = pre_c;
  }
  if (!COMPRESS && !PARK && threadIdx.x == 0) tma_store_wait_all();  // shared memory must outlive the last bulk store
}
#endif
