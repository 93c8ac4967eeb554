// scan.cu — hand-written prefix-sum and stream-compaction kernels for sm_100a: single-pass decoupled look-back,
// persistent CTAs, TMA bulk copies in (and, for prefix sums, out).  HBM-bound integer / byte movers: what matters is
// keeping the memory system busy while the tile-to-tile prefix chain makes progress (see the history in
// profiles/r01_scan_history.md).  No tensor cores on purpose.
//
// There is no reference implementation of these operations; the CPU oracle (oracle/oracle.cpp) is the
// specification they are tested against.
#include "prims.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <type_traits>
#include <vector>

#include "common.h"

namespace vkjit {
namespace prims {

#include "scan_common.cuh"  // streaming ld/st, tile status words + look-back, TMA/mbarrier (shared with the NVRTC kernels)

// ---------------------------------------------------------------------------------------
// decoupled look-back scan (Merrill & Garland) — prefix sum and stream compaction
// ---------------------------------------------------------------------------------------
// Tile = 1024 threads x VPT vectors x 4 lanes (24576 lanes = 96 KiB for the scans, see ScanGeom).  Vector
// q = j*1024 + t of a tile is held by thread t in register slot j, so every shared-memory read
// and every global store of a warp covers 512 contiguous bytes.  Tile status words pack
// {flag:32 | value:32} into one 64-bit word so flag and value travel in a single (relaxed,
// L2-coherent) access and no fence is needed between them.
// Why the tile is this large: every generation of 148 tiles pays one cross-SM aggregate exchange, and
// with strided persistent tiles each generation runs at the pace of its slowest SM; fewer, larger
// generations pay that less often (history: profiles/r01_scan_history.md).
constexpr uint32_t kValueStageDensity = 16;  // compress -> values: TMA-stage the values of tiles selecting >= 1/16 of their lanes
constexpr int kCompressLagVptIndex = 4, kCompressLagVptValues = 3;  // lagged compaction: 16384- / 12288-lane tiles
constexpr int kScanLagVpt = 4;  // lagged variant: 16384-lane tiles, 3 ring slots
enum ScanMode { MODE_EXCLUSIVE = 0, MODE_INCLUSIVE = 1, MODE_COMPRESS_INDEX = 2, MODE_COMPRESS_VALUE = 3 };

// Persistent kernel: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...  The input
// tiles arrive through a 2-stage shared-memory ring filled by TMA bulk copies, so up to
// 2 x 96 KiB of loads per SM stay in flight while the CTA is scanning, waiting at a
// barrier or looking back — the three serial phases no longer starve the memory system
// (the first version, 2 non-persistent CTAs/SM with register loads, spent 60 % of its stall
// samples at barriers and reached 45 % DRAM utilisation).
// Tile geometry.  Fewer, larger generations amortise the per-generation look-back cost (measured at 2^28:
// prefix sum 0.46 ms with 16384-lane tiles / 3 stages, 0.39 ms with 24576-lane tiles / 2 stages; compress
// 0.65 -> 0.59 ms).  1024 threads x 6 vectors is the largest tile that stays within 64 registers.
template <int MODE> struct ScanGeom {
  static constexpr int TILE = 24576;
  static constexpr int STAGES = 2;
  static constexpr size_t SMEM = (size_t)STAGES * TILE * 4;
};

template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel(const uint32_t* __restrict__ in,      // scan: addends; compress: mask words
            const uint32_t* __restrict__ values,  // MODE_COMPRESS_VALUE only
            uint32_t* __restrict__ out, uint32_t* __restrict__ count_out, size_t n, uint32_t num_tiles,
            uint64_t* __restrict__ state, uint32_t diag_skip_lookback, const uint32_t* __restrict__ initial_ptr,
            const uint32_t* __restrict__ index_base_ptr) {  // MODE_COMPRESS_INDEX: global index of lane 0 (sharded masks)
  constexpr int T = kScanThreads;
  constexpr int kScanTile = ScanGeom<MODE>::TILE;
  constexpr int kScanStages = ScanGeom<MODE>::STAGES;
  constexpr int VPT = kScanTile / (T * 4);  // 128-bit vectors per thread
  constexpr int WARPS = T / 32;
  constexpr int NTOT = VPT * WARPS;         // (slot, warp) totals per tile
  constexpr int PER_LANE = NTOT / 32;
  constexpr int S = kScanStages;
  constexpr uint32_t TILE_BYTES = kScanTile * 4;
  constexpr bool COMPRESS = MODE >= MODE_COMPRESS_INDEX;
  static_assert(NTOT % 32 == 0, "tile totals must fill whole warp rows");
  extern __shared__ __align__(128) unsigned char ring_raw[];
  uint32_t* ring = reinterpret_cast<uint32_t*>(ring_raw);  // S stages x kScanTile words
  __shared__ __align__(8) uint64_t full[S];
  __shared__ uint32_t s_tot[2][NTOT];
  __shared__ uint32_t s_tile_excl[2];
  __shared__ uint32_t s_vstaged[2];
  // compress -> values: when a tile selects enough lanes, its VALUES are fetched by TMA into the ring slot its mask
  // words just left, during the look-back, instead of by register loads after it (whose latency is exposed once
  // per tile).  The slot's barrier then completes twice for that tile, so its parity is tracked, not derived.
  constexpr bool VSTAGE = MODE == MODE_COMPRESS_VALUE;
  uint32_t parity = 0;  // bit s: parity of the next completion of full[s]

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t first = blockIdx.x, stride = gridDim.x;
  const uint32_t my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % kScanTile) != 0;  // the globally last tile is partial: plain guarded loads
  const uint32_t index_base = (MODE == MODE_COMPRESS_INDEX && index_base_ptr) ? __ldcg(index_base_ptr) : 0u;

  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (uint32_t k = 0; k < (uint32_t)S && k < my_tiles; ++k) {
      const uint32_t t = first + k * stride;
      if (ragged && t == num_tiles - 1) continue;
      mbar_expect_tx(&full[k], TILE_BYTES);
      tma_load_1d(ring + (size_t)k * kScanTile, in + (size_t)t * kScanTile, TILE_BYTES, &full[k]);
    }
  }

  for (uint32_t k = 0; k < my_tiles; ++k) {
    const uint32_t tile = first + k * stride;
    const int stage = k % S;
    const int buf = k & 1;
    const size_t tile_base = (size_t)tile * kScanTile;
    const bool staged = !(ragged && tile == num_tiles - 1);

    uint4 x[VPT];  // scan: addends; compress: mask words
    if (staged) {
      mbar_wait(&full[stage], (parity >> stage) & 1u);
      parity ^= 1u << stage;
      const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)stage * kScanTile);
#pragma unroll
      for (int j = 0; j < VPT; ++j) x[j] = src[j * T + threadIdx.x];  // conflict-free 128-bit shared loads
    } else {
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
        if (e + 3 < n) x[j] = ld_stream(reinterpret_cast<const uint4*>(in + e));
        else {
          x[j].x = e + 0 < n ? in[e + 0] : 0u; x[j].y = e + 1 < n ? in[e + 1] : 0u;
          x[j].z = e + 2 < n ? in[e + 2] : 0u; x[j].w = 0u;
        }
      }
    }

    // 1) per-vector sums, 2) inclusive warp scan per register slot, 3) one warp scans the
    // (slot, warp) totals in tile order, 4) look-back gives the tile's global offset.
    // The kernel is as much instruction-issue bound as memory bound (ncu: ~830 warp instructions per warp and
    // tile, issue slots 50 % busy at half occupancy), so the per-lane instruction count is kept down.
    uint32_t flags[VPT];  // compress: 4 selection bits per vector
    uint32_t vsum[VPT], wincl[VPT];
    if (COMPRESS) {
      // a warp's inclusive count per slot is at most 32 x 4 = 128: four slots share one 32-bit word (8-bit
      // fields, no carries between them), so VPT warp scans become ceil(VPT / 4)
      constexpr int G = (VPT + 3) / 4;
      uint32_t pk[G];
#pragma unroll
      for (int g = 0; g < G; ++g) pk[g] = 0u;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        flags[j] = (x[j].x != 0u ? 1u : 0u) | (x[j].y != 0u ? 2u : 0u) | (x[j].z != 0u ? 4u : 0u) | (x[j].w != 0u ? 8u : 0u);
        vsum[j] = (uint32_t)__popc(flags[j]);
        pk[j / 4] |= vsum[j] << (8 * (j % 4));
      }
#pragma unroll
      for (int g = 0; g < G; ++g) {
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pk[g], o);
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
        uint32_t s = vsum[j];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += t;
        }
        wincl[j] = s;
        if (lane == 31) s_tot[buf][j * WARPS + warp] = s;
      }
    }
    __syncthreads();  // every thread has consumed its part of ring[stage]: the stage can be refilled
    auto refill = [&]() {  // thread 0: next mask / addend tile of this slot
      if (k + S < my_tiles) {
        const uint32_t t2 = first + (k + S) * stride;
        if (!(ragged && t2 == num_tiles - 1)) {
          mbar_expect_tx(&full[stage], TILE_BYTES);
          tma_load_1d(ring + (size_t)stage * kScanTile, in + (size_t)t2 * kScanTile, TILE_BYTES, &full[stage]);
        }
      }
    };
    if (!VSTAGE && threadIdx.x == 32) refill();  // warp 1 issues the copy: a full TMA queue would otherwise hold up warp 0's look-back
    if (warp == 0) {
      uint32_t t[PER_LANE], run = 0;
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) { t[i] = s_tot[buf][lane * PER_LANE + i]; run += t[i]; }
      uint32_t s = run;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, s, o);
        if (lane >= o) s += u;
      }
      uint32_t off = s - run;  // exclusive offset of this lane's first entry
#pragma unroll
      for (int i = 0; i < PER_LANE; ++i) { s_tot[buf][lane * PER_LANE + i] = off; off += t[i]; }
      const uint32_t aggregate = __shfl_sync(0xFFFFFFFFu, s, 31);
      if (VSTAGE && lane == 0) {
        const bool vs = staged && aggregate * kValueStageDensity >= (uint32_t)kScanTile;
        s_vstaged[buf] = vs ? 1u : 0u;
        if (vs) {
          mbar_expect_tx(&full[stage], TILE_BYTES);
          tma_load_1d(ring + (size_t)stage * kScanTile, values + tile_base, TILE_BYTES, &full[stage]);
        } else {
          refill();
        }
      }
      // diag_skip_lookback: timing-only diagnostic (VKJIT_SCAN_DIAG=nolookback), results are wrong
      const uint32_t initial = (tile == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
      const uint32_t excl = diag_skip_lookback ? 0u : look_back(status, tile, aggregate, initial);
      if (lane == 0) {
        s_tile_excl[buf] = excl;
        if (COMPRESS && tile == num_tiles - 1) *count_out = excl + aggregate;
      }
    }
    __syncthreads();
    const uint32_t tile_excl = s_tile_excl[buf];
    const bool vstaged = VSTAGE && s_vstaged[buf] != 0u;
    if (vstaged) {
      mbar_wait(&full[stage], (parity >> stage) & 1u);
      parity ^= 1u << stage;
    }
    const uint4* vsrc = reinterpret_cast<const uint4*>(ring + (size_t)stage * kScanTile);

    // W: whole tile — the common case is compiled without bounds checks
    auto emit = [&](auto whole_c) {
      constexpr bool W = decltype(whole_c)::value;
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
        const uint32_t p = tile_excl + s_tot[buf][j * WARPS + warp] + (wincl[j] - vsum[j]);  // exclusive prefix of lane e
        if (!COMPRESS) {
          uint4 r;
          if (MODE == MODE_EXCLUSIVE) { r.x = p; r.y = p + x[j].x; r.z = r.y + x[j].y; r.w = r.z + x[j].z; }
          else { r.x = p + x[j].x; r.y = r.x + x[j].y; r.z = r.y + x[j].z; r.w = r.z + x[j].w; }
          if (W || e + 3 < n) st_stream(reinterpret_cast<uint4*>(out + e), r);
          else {
            if (e + 0 < n) out[e + 0] = r.x;
            if (e + 1 < n) out[e + 1] = r.y;
            if (e + 2 < n) out[e + 2] = r.z;
          }
        } else if (flags[j]) {
          // selected lanes are written at their rank; flags of out-of-range lanes are 0
          const uint32_t f = flags[j];
          uint4 v;
          if (MODE == MODE_COMPRESS_VALUE) {  // values are read once, only for vectors with a selected lane
            if (W && vstaged) v = vsrc[j * T + threadIdx.x];
            else if (W || e + 3 < n) v = ld_stream(reinterpret_cast<const uint4*>(values + e));
            else {
              v.x = e + 0 < n ? values[e + 0] : 0u; v.y = e + 1 < n ? values[e + 1] : 0u;
              v.z = e + 2 < n ? values[e + 2] : 0u; v.w = 0u;
            }
          } else { v.x = index_base + (uint32_t)e; v.y = v.x + 1; v.z = v.x + 2; v.w = v.x + 3; }
          uint32_t* q = out + p;  // one 64-bit address per vector; the slots of its lanes follow from the flag bits
          const uint32_t s1 = f & 1u, s2 = s1 + ((f >> 1) & 1u), s3 = s2 + ((f >> 2) & 1u);
          if (f & 1u) q[0] = v.x;
          if (f & 2u) q[s1] = v.y;
          if (f & 4u) q[s2] = v.z;
          if (f & 8u) q[s3] = v.w;
        }
      }
    };
    if (staged) emit(std::true_type{}); else emit(std::false_type{});
    if (vstaged) {  // the slot held the values until now
      __syncthreads();
      if (threadIdx.x == 0) refill();
    }
    // s_tot/s_tile_excl are double-buffered: iteration k+2 rewrites buffer `buf` only after every
    // thread passed the first barrier of iteration k+1, i.e. after it finished reading it here.
  }
}

// Lagged prefix sum (the default for MODE_EXCLUSIVE / MODE_INCLUSIVE; VKJIT_SCAN_IMPL=classic selects scan_kernel).
// Measured on the kernel above (globaltimer per phase, profiles/): a tile
// costs 0.9 us local scan + 2.2-2.5 us look-back + 1.4 us stores, and the look-back time is the same at every
// position of a generation — it is not a wait for stragglers but the store -> poll -> load visibility latency of
// the immediate predecessor's status word, which the 148 CTAs chase around a ring.  Here tile k's aggregate is
// published as soon as its local scan is done, but its prefix is resolved one iteration later (by then every
// predecessor has been visible for a whole tile period: one L2 round trip), and its results are written from
// registers then.  All three ring slots stay free for prefetching.
template <int MODE>
__global__ void __launch_bounds__(kScanThreads, 1)
scan_kernel_lag(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, size_t n, uint32_t num_tiles,
                uint64_t* __restrict__ state, const uint32_t* __restrict__ initial_ptr, unsigned long long* __restrict__ trace,
                const bool early) {  // early: request a tile's status window right after the previous resolve ($VKJIT_SCAN_EARLY_PRIM)
  constexpr int T = kScanThreads, VPT = kScanLagVpt, TILE = T * 4 * VPT, S = 2;
  constexpr int WARPS = T / 32, NTOT = VPT * WARPS, PER_LANE = NTOT / 32;
  constexpr uint32_t TILE_BYTES = TILE * 4;
  static_assert(NTOT % 32 == 0, "tile totals must fill whole warp rows");
  extern __shared__ __align__(128) unsigned char ring_raw[];
  uint32_t* ring = reinterpret_cast<uint32_t*>(ring_raw);  // S input slots, then one output staging tile
  uint32_t* stage_out = ring + (size_t)S * TILE;
  __shared__ __align__(8) uint64_t full[S];
  __shared__ uint32_t s_tot[3][NTOT];  // tile k writes [k % 3] while tile k-1's offsets are still being read
  __shared__ uint32_t s_tile_excl;
  __shared__ __align__(16) uint64_t s_window[kLookWide * 32 * 2];

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t first = blockIdx.x, stride = gridDim.x;
  const uint32_t my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % TILE) != 0;

  auto fill = [&](uint32_t k) {  // thread 0
    if (k >= my_tiles) return;
    const uint32_t t = first + k * stride;
    if (ragged && t == num_tiles - 1) return;
    mbar_expect_tx(&full[k % S], TILE_BYTES);
    tma_load_1d(ring + (size_t)(k % S) * TILE, in + (size_t)t * TILE, TILE_BYTES, &full[k % S]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (uint32_t k = 0; k < (uint32_t)S; ++k) fill(k);

  uint4 xp[VPT];          // previous tile: results relative to the start of the vector's warp row
  uint32_t agg_prev = 0u; // warp 0: aggregate of the previous tile
#pragma unroll
  for (int j = 0; j < VPT; ++j) xp[j] = make_uint4(0u, 0u, 0u, 0u);

  for (uint32_t k = 0; k <= my_tiles; ++k) {
    const bool have_cur = k < my_tiles, have_prev = k > 0;
    const uint32_t tile = first + k * stride;
    uint4 xc[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) xc[j] = make_uint4(0u, 0u, 0u, 0u);
    // warp 0: the previous tile's predecessors published a whole iteration ago — fetch their status words now,
    // use them after this tile's local scan
    if (!early && warp == 0 && have_prev) prefetch_window(status, tile - stride, s_window);
    if (have_cur) {  // ---- local scan of tile k
      const size_t tile_base = (size_t)tile * TILE;
      if (trace && threadIdx.x == 0) trace[(size_t)tile * 8 + 0] = global_ns();
      if (!(ragged && tile == num_tiles - 1)) {
        mbar_wait(&full[k % S], (k / S) & 1);
        const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)(k % S) * TILE);
#pragma unroll
        for (int j = 0; j < VPT; ++j) xc[j] = src[j * T + threadIdx.x];
      } else {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          if (e + 3 < n) xc[j] = ld_stream(reinterpret_cast<const uint4*>(in + e));
          else {
            xc[j].x = e + 0 < n ? in[e + 0] : 0u; xc[j].y = e + 1 < n ? in[e + 1] : 0u;
            xc[j].z = e + 2 < n ? in[e + 2] : 0u; xc[j].w = 0u;
          }
        }
      }
      if (trace && threadIdx.x == 0) trace[(size_t)tile * 8 + 1] = global_ns();
      uint32_t* tot = s_tot[k % 3];
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const uint32_t vs = xc[j].x + xc[j].y + xc[j].z + xc[j].w;
        uint32_t s = vs;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += t;
        }
        if (lane == 31) tot[j * WARPS + warp] = s;
        const uint32_t p = s - vs;  // exclusive offset of the vector within its warp row
        const uint4 a = xc[j];
        if (MODE == MODE_EXCLUSIVE) { xc[j].x = p; xc[j].y = p + a.x; xc[j].z = xc[j].y + a.y; xc[j].w = xc[j].z + a.z; }
        else { xc[j].x = p + a.x; xc[j].y = xc[j].x + a.y; xc[j].z = xc[j].y + a.z; xc[j].w = xc[j].z + a.w; }
      }
    }
    __syncthreads();  // s_tot[k % 3] complete; every thread has consumed ring slot k % S
    if (have_cur && threadIdx.x == 32) fill(k + S);
    if (have_cur && trace && threadIdx.x == 0) trace[(size_t)tile * 8 + 2] = global_ns();
    if (warp == 0) {
      uint32_t agg_cur = 0u;
      if (have_cur) {  // (slot, warp) totals in tile order -> exclusive offsets; publish the tile aggregate
        uint32_t* tot = s_tot[k % 3];
        uint32_t t[PER_LANE], run = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { t[i] = tot[lane * PER_LANE + i]; run += t[i]; }
        uint32_t s = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += u;
        }
        uint32_t off = s - run;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { tot[lane * PER_LANE + i] = off; off += t[i]; }
        agg_cur = __shfl_sync(0xFFFFFFFFu, s, 31);
        const uint32_t initial = (tile == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u;
        if (lane == 0) {
          if (tile == 0) status_store(status, ((uint64_t)ST_INCLUSIVE << 32) | (initial + agg_cur));
          else status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
        }
      }
      if (have_prev) {  // the previous tile's predecessors have been visible for a whole iteration
        const uint32_t tprev = tile - stride;
        if (trace && lane == 0) trace[(size_t)tprev * 8 + 6] = global_ns();
        uint64_t window[kLookWide];
        load_window(s_window, window);
        const uint32_t excl = resolve_prefix(status, tprev, agg_prev, (tprev == 0 && initial_ptr) ? __ldcg(initial_ptr) : 0u, window);
        if (lane == 0) {
          s_tile_excl = excl;
          if (trace) { trace[(size_t)tprev * 8 + 3] = global_ns(); trace[(size_t)tprev * 8 + 5] = blockIdx.x; }
        }
      }
      agg_prev = agg_cur;
    }
    if (threadIdx.x == 0) tma_store_wait_read();  // the previous bulk store has read the staging tile
    __syncthreads();
    if (early && warp == 0 && have_cur) prefetch_window(status, tile, s_window);  // resolved in the next iteration
    if (have_prev) {  // ---- results of tile k-1, from registers, through shared memory and ONE bulk store
      const uint32_t tprev = tile - stride;
      const size_t tile_base = (size_t)tprev * TILE;
      const bool whole = !(ragged && tprev == num_tiles - 1);
      const uint32_t tile_excl = s_tile_excl;
      const uint32_t* tot = s_tot[(k - 1) % 3];
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
        const uint32_t p = tile_excl + tot[j * WARPS + warp];
        uint4 r;
        r.x = xp[j].x + p; r.y = xp[j].y + p; r.z = xp[j].z + p; r.w = xp[j].w + p;
        if (whole) reinterpret_cast<uint4*>(stage_out)[j * T + threadIdx.x] = r;
        else if (e + 3 < n) st_stream(reinterpret_cast<uint4*>(out + e), r);
        else {
          if (e + 0 < n) out[e + 0] = r.x;
          if (e + 1 < n) out[e + 1] = r.y;
          if (e + 2 < n) out[e + 2] = r.z;
        }
      }
      if (whole) {
        fence_proxy_async();
        __syncthreads();
        if (threadIdx.x == 0) tma_store_1d(out + tile_base, stage_out, TILE_BYTES);
      }
      if (trace && threadIdx.x == 0) trace[(size_t)tprev * 8 + 4] = global_ns();
    }
#pragma unroll
    for (int j = 0; j < VPT; ++j) xp[j] = xc[j];
  }
  if (threadIdx.x == 0) tma_store_wait_all();  // shared memory must outlive the last bulk store
}

// Lagged stream compaction (the default for MODE_COMPRESS_*; VKJIT_SCAN_IMPL=classic selects scan_kernel): the same
// schedule as scan_kernel_lag.  What a tile carries into the next iteration is tiny — 4 selection bits and one
// 8-bit row offset per vector, packed into two registers — so nothing is parked in shared memory.
//   VALUES = false: 16384-lane tiles, 3 mask slots.
//   VALUES = true : 12288-lane tiles, 2 mask slots + 2 value slots; the values of tile k are fetched by one TMA
//                   bulk copy issued as soon as its count is known (dense tiles only) and read from shared memory
//                   when the tile is written one iteration later.
template <bool VALUES>
__global__ void __launch_bounds__(kScanThreads, 1)
compress_kernel_lag(const uint32_t* __restrict__ mask, const uint32_t* __restrict__ values, uint32_t* __restrict__ out,
                    uint32_t* __restrict__ count_out, size_t n, uint32_t num_tiles, uint64_t* __restrict__ state,
                    const uint32_t* __restrict__ index_base_ptr, const bool early) {
  constexpr int T = kScanThreads, VPT = VALUES ? kCompressLagVptValues : kCompressLagVptIndex, TILE = T * 4 * VPT;
  constexpr int S = VALUES ? 2 : 3;  // mask slots
  constexpr int WARPS = T / 32, NTOT = VPT * WARPS, PER_LANE = NTOT / 32;
  constexpr uint32_t TILE_BYTES = TILE * 4;
  static_assert(NTOT % 32 == 0 && VPT <= 4, "packed per-slot counts need VPT <= 4");
  extern __shared__ __align__(128) unsigned char ring_raw[];
  uint32_t* ring = reinterpret_cast<uint32_t*>(ring_raw);  // S mask slots, then (VALUES) 2 value slots
  uint32_t* vring = ring + (size_t)S * TILE;
  __shared__ __align__(8) uint64_t full[S];
  __shared__ __align__(8) uint64_t vfull[2];
  __shared__ uint32_t s_tot[3][NTOT];
  __shared__ uint32_t s_tile_excl;
  __shared__ uint32_t s_vstaged[2];
  __shared__ __align__(16) uint64_t s_window[kLookWide * 32 * 2];

  uint64_t* status = state + kStatusStride;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t first = blockIdx.x, stride = gridDim.x;
  const uint32_t my_tiles = first < num_tiles ? (num_tiles - first + stride - 1) / stride : 0;
  const bool ragged = (n % TILE) != 0;
  const uint32_t index_base = (!VALUES && index_base_ptr) ? __ldcg(index_base_ptr) : 0u;

  auto fill = [&](uint32_t k) {  // next mask tile of slot k % S
    if (k >= my_tiles) return;
    const uint32_t t = first + k * stride;
    if (ragged && t == num_tiles - 1) return;
    mbar_expect_tx(&full[k % S], TILE_BYTES);
    tma_load_1d(ring + (size_t)(k % S) * TILE, mask + (size_t)t * TILE, TILE_BYTES, &full[k % S]);
  };
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) mbar_init(&full[s], 1);
    mbar_init(&vfull[0], 1); mbar_init(&vfull[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x == 0)
    for (uint32_t k = 0; k < (uint32_t)S; ++k) fill(k);

  uint32_t flags_p = 0u, pre_p = 0u;  // previous tile: 4 selection bits / one 8-bit exclusive row offset per vector
  uint32_t agg_prev = 0u;             // warp 0: aggregate of the previous tile
  uint32_t vparity = 0u;              // bit s: parity of the next completion of vfull[s]

  for (uint32_t k = 0; k <= my_tiles; ++k) {
    const bool have_cur = k < my_tiles, have_prev = k > 0;
    const uint32_t tile = first + k * stride;
    const bool staged = have_cur && !(ragged && tile == num_tiles - 1);
    uint32_t flags_c = 0u, pre_c = 0u;
    if (!early && warp == 0 && have_prev) prefetch_window(status, tile - stride, s_window);
    if (have_cur) {  // ---- selection bits and packed local scan of tile k
      const size_t tile_base = (size_t)tile * TILE;
      uint32_t pk = 0u, own = 0u;
      if (staged) {
        mbar_wait(&full[k % S], (k / S) & 1);
        const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)(k % S) * TILE);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const uint4 x = src[j * T + threadIdx.x];
          const uint32_t f = (x.x != 0u ? 1u : 0u) | (x.y != 0u ? 2u : 0u) | (x.z != 0u ? 4u : 0u) | (x.w != 0u ? 8u : 0u);
          flags_c |= f << (4 * j);
          own |= (uint32_t)__popc(f) << (8 * j);
        }
      } else {
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          uint32_t f = 0u;
#pragma unroll
          for (int c = 0; c < 4; ++c) f |= (e + c < n && mask[e + c] != 0u) ? (1u << c) : 0u;
          flags_c |= f << (4 * j);
          own |= (uint32_t)__popc(f) << (8 * j);
        }
      }
      pk = own;  // a warp's inclusive count per slot is at most 128: four 8-bit fields scan in one word
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, pk, o);
        if (lane >= o) pk += t;
      }
      pre_c = pk - own;
      if (lane == 31) {
#pragma unroll
        for (int j = 0; j < VPT; ++j) s_tot[k % 3][j * WARPS + warp] = (pk >> (8 * j)) & 0xFFu;
      }
    }
    __syncthreads();  // s_tot[k % 3] complete; every thread has consumed mask slot k % S
    if (have_cur && threadIdx.x == 32) fill(k + S);
    if (warp == 0) {
      uint32_t agg_cur = 0u;
      if (have_cur) {
        uint32_t* tot = s_tot[k % 3];
        uint32_t t[PER_LANE], run = 0;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { t[i] = tot[lane * PER_LANE + i]; run += t[i]; }
        uint32_t s = run;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, s, o);
          if (lane >= o) s += u;
        }
        uint32_t off = s - run;
#pragma unroll
        for (int i = 0; i < PER_LANE; ++i) { tot[lane * PER_LANE + i] = off; off += t[i]; }
        agg_cur = __shfl_sync(0xFFFFFFFFu, s, 31);
        if (lane == 0) {
          if (tile == 0) status_store(status, ((uint64_t)ST_INCLUSIVE << 32) | agg_cur);
          else status_store(status + (size_t)tile * kStatusStride, ((uint64_t)ST_AGGREGATE << 32) | agg_cur);
          if (VALUES) {  // dense tile: its values travel by TMA while the next tile is scanned
            const bool vs = staged && agg_cur * kValueStageDensity >= (uint32_t)TILE;
            s_vstaged[k % 2] = vs ? 1u : 0u;
            if (vs) {
              mbar_expect_tx(&vfull[k % 2], TILE_BYTES);
              tma_load_1d(vring + (size_t)(k % 2) * TILE, values + (size_t)tile * TILE, TILE_BYTES, &vfull[k % 2]);
            }
          }
        }
      }
      if (have_prev) {
        const uint32_t tprev = tile - stride;
        uint64_t window[kLookWide];
        load_window(s_window, window);
        const uint32_t excl = resolve_prefix(status, tprev, agg_prev, 0u, window);
        if (lane == 0) {
          s_tile_excl = excl;
          if (tprev == num_tiles - 1) *count_out = excl + agg_prev;
        }
      }
      agg_prev = agg_cur;
    }
    __syncthreads();
    if (early && warp == 0 && have_cur) prefetch_window(status, tile, s_window);  // resolved in the next iteration
    if (have_prev) {  // ---- selected lanes of tile k-1, written at their rank
      const uint32_t kp = k - 1, tprev = tile - stride;
      const size_t tile_base = (size_t)tprev * TILE;
      const bool whole = !(ragged && tprev == num_tiles - 1);
      const uint32_t tile_excl = s_tile_excl;
      const uint32_t* tot = s_tot[kp % 3];
      bool vstaged = false;
      if (VALUES) {
        vstaged = s_vstaged[kp % 2] != 0u;
        if (vstaged) {
          mbar_wait(&vfull[kp % 2], (vparity >> (kp % 2)) & 1u);
          vparity ^= 1u << (kp % 2);
        }
      }
      const uint4* vsrc = reinterpret_cast<const uint4*>(vring + (size_t)(kp % 2) * TILE);
#pragma unroll
      for (int j = 0; j < VPT; ++j) {
        const uint32_t f = (flags_p >> (4 * j)) & 0xFu;
        if (f) {
          const size_t e = tile_base + ((size_t)j * T + threadIdx.x) * 4;
          uint4 v;
          if (VALUES) {
            if (vstaged) v = vsrc[j * T + threadIdx.x];
            else if (whole || e + 3 < n) v = ld_stream(reinterpret_cast<const uint4*>(values + e));
            else {
              v.x = e + 0 < n ? values[e + 0] : 0u; v.y = e + 1 < n ? values[e + 1] : 0u;
              v.z = e + 2 < n ? values[e + 2] : 0u; v.w = 0u;
            }
          } else { v.x = index_base + (uint32_t)e; v.y = v.x + 1; v.z = v.x + 2; v.w = v.x + 3; }
          uint32_t* q = out + (tile_excl + tot[j * WARPS + warp] + ((pre_p >> (8 * j)) & 0xFFu));
          const uint32_t s1 = f & 1u, s2 = s1 + ((f >> 1) & 1u), s3 = s2 + ((f >> 2) & 1u);
          if (f & 1u) q[0] = v.x;
          if (f & 2u) q[s1] = v.y;
          if (f & 4u) q[s2] = v.z;
          if (f & 8u) q[s3] = v.w;
        }
      }
    }
    flags_p = flags_c; pre_p = pre_c;
  }
}

static_assert(kStatusWordsPerTile == kStatusStride, "host and device disagree on the status slot size");
size_t scan_state_words(size_t n, size_t tile) { return (size_t)kStatusStride * (2 + n / tile); }

template <int MODE>
static void launch_scan(const uint32_t* in, const uint32_t* values, uint32_t* out, uint32_t* count_out, size_t n,
                        const Scratch& sc, int sm_count, cudaStream_t s, const uint32_t* initial = nullptr,
                        const uint32_t* index_base = nullptr) {
  using G = ScanGeom<MODE>;
  static int impl = -1;
  if (impl < 0) { const char* d = getenv("VKJIT_SCAN_IMPL"); impl = (d && std::string(d) == "classic") ? 0 : 1; }  // prefix sums: lagged kernel unless "classic"
  static const bool early = [] { const char* e = getenv("VKJIT_SCAN_EARLY_PRIM"); return e && e[0] == '1'; }();  // A/B knob (default: off)
  if constexpr (MODE < MODE_COMPRESS_INDEX) {
    if (impl == 1) {
      constexpr size_t TILE = (size_t)kScanThreads * 4 * kScanLagVpt, SMEM = 3 * TILE * 4;  // 2 input slots + 1 output staging tile
      const size_t tiles = (n + TILE - 1) / TILE;
      const size_t words = (size_t)kStatusStride * (1 + tiles);
      if (words > sc.tile_state_words) fail(VKJIT_ERR_INVALID, "scan scratch too small");
      cudaError_t e = cudaMemsetAsync(sc.tile_state, 0, words * sizeof(uint64_t), s);
      if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan memset: ") + cudaGetErrorString(e));
      static bool configured_lag = false;
      if (!configured_lag) {
        e = cudaFuncSetAttribute(scan_kernel_lag<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan smem attribute: ") + cudaGetErrorString(e));
        configured_lag = true;
      }
      static unsigned long long* d_trace = nullptr;
      const char* tf = getenv("VKJIT_SCAN_TRACE");
      if (tf && !d_trace) cudaMalloc(&d_trace, tiles * 64);
      if (tf) cudaMemsetAsync(d_trace, 0, tiles * 64, s);
      const unsigned grid = (unsigned)std::min<size_t>(tiles, (size_t)sm_count);
      scan_kernel_lag<MODE><<<grid, kScanThreads, SMEM, s>>>(in, out, n, (uint32_t)tiles, sc.tile_state, initial, tf ? d_trace : nullptr, early);
      e = cudaGetLastError();
      if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan launch: ") + cudaGetErrorString(e));
      if (tf) {
        cudaStreamSynchronize(s);
        std::vector<unsigned long long> h(tiles * 8);
        cudaMemcpy(h.data(), d_trace, tiles * 64, cudaMemcpyDeviceToHost);
        FILE* fp = fopen(tf, "wb");
        if (fp) { fwrite(h.data(), 8, h.size(), fp); fclose(fp); }
      }
      return;
    }
  }
  if constexpr (MODE >= MODE_COMPRESS_INDEX) {
    if (impl == 1) {
      constexpr bool VALUES = MODE == MODE_COMPRESS_VALUE;
      constexpr size_t TILE = (size_t)kScanThreads * 4 * (VALUES ? kCompressLagVptValues : kCompressLagVptIndex);
      constexpr size_t SMEM = (VALUES ? 4 : 3) * TILE * 4;
      const size_t tiles = (n + TILE - 1) / TILE;
      const size_t words = (size_t)kStatusStride * (1 + tiles);
      if (words > sc.tile_state_words) fail(VKJIT_ERR_INVALID, "scan scratch too small");
      cudaError_t e = cudaMemsetAsync(sc.tile_state, 0, words * sizeof(uint64_t), s);
      if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan memset: ") + cudaGetErrorString(e));
      static bool configured_lag = false;
      if (!configured_lag) {
        e = cudaFuncSetAttribute(compress_kernel_lag<VALUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM);
        if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("compress smem attribute: ") + cudaGetErrorString(e));
        configured_lag = true;
      }
      const unsigned grid = (unsigned)std::min<size_t>(tiles, (size_t)sm_count);
      compress_kernel_lag<VALUES><<<grid, kScanThreads, SMEM, s>>>(in, values, out, count_out, n, (uint32_t)tiles, sc.tile_state, index_base, early);
      e = cudaGetLastError();
      if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("compress launch: ") + cudaGetErrorString(e));
      return;
    }
  }
  const size_t tiles = (n + G::TILE - 1) / G::TILE;
  const size_t words = (size_t)kStatusStride * (1 + tiles);
  if (words > sc.tile_state_words) fail(VKJIT_ERR_INVALID, "scan scratch too small");
  cudaError_t e = cudaMemsetAsync(sc.tile_state, 0, words * sizeof(uint64_t), s);
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan memset: ") + cudaGetErrorString(e));
  static bool configured = false;
  if (!configured) {
    e = cudaFuncSetAttribute(scan_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan smem attribute: ") + cudaGetErrorString(e));
    configured = true;
  }
  // one persistent CTA per SM: all CTAs are co-resident, so a tile only ever waits on tiles of
  // CTAs that are running (forward progress of the look-back does not depend on dispatch order)
  const unsigned grid = (unsigned)std::min<size_t>(tiles, (size_t)sm_count);
  static int diag = -1;
  if (diag < 0) { const char* d = getenv("VKJIT_SCAN_DIAG"); diag = (d && std::string(d) == "nolookback") ? 1 : 0; }
  scan_kernel<MODE><<<grid, kScanThreads, G::SMEM, s>>>(in, values, out, count_out, n, (uint32_t)tiles, sc.tile_state, (uint32_t)diag, initial, index_base);
  e = cudaGetLastError();
  if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("scan launch: ") + cudaGetErrorString(e));
}

void prefix_sum(const uint32_t* in, uint32_t* out, size_t n, bool exclusive, const Scratch& sc, int sm_count, void* stream,
                const uint32_t* initial) {
  if (n == 0) return;
  cudaStream_t s = (cudaStream_t)stream;
  if (exclusive) launch_scan<MODE_EXCLUSIVE>(in, nullptr, out, nullptr, n, sc, sm_count, s, initial);
  else launch_scan<MODE_INCLUSIVE>(in, nullptr, out, nullptr, n, sc, sm_count, s, initial);
}

void compress(const uint32_t* mask, const uint32_t* values, uint32_t* out, uint32_t* count_out, size_t n,
              const Scratch& sc, int sm_count, void* stream, const uint32_t* index_base) {
  if (n == 0) return;
  cudaStream_t s = (cudaStream_t)stream;
  if (values) launch_scan<MODE_COMPRESS_VALUE>(mask, values, out, count_out, n, sc, sm_count, s);
  else launch_scan<MODE_COMPRESS_INDEX>(mask, nullptr, out, count_out, n, sc, sm_count, s, nullptr, index_base);
}

}  // namespace prims
}  // namespace vkjit
