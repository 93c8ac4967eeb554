// program.h — canonical, address-free form of a scheduled trace + CUDA C generation.
//
// Replaces the reference's `Kernel` (libs/vkjit-core/src/internal.rs:557-1314), which bakes
// buffer addresses into the SPIR-V as 64-bit constants (internal.rs:614-633) and therefore
// can never be cached.  Here a trace is walked in the order `record_ops` visits it
// (internal.rs:874-1117), vars are renumbered 0..k, and everything that varies between
// launches of the same computation (addresses, n, the arange count, the shard base) becomes
// a kernel parameter.  The canonical word string is the cache key.
#pragma once
#include <string>
#include <vector>

#include "ir.h"

namespace vkjit {

struct Hash128 {
  uint64_t lo = 0, hi = 0;
  bool operator==(const Hash128& o) const { return lo == o.lo && hi == o.hi; }
};

enum ParamUse : uint8_t { USE_STREAM = 1, USE_GATHER = 2, USE_SCATTER = 4 };

// fused trace -> scan kernels (scan_fused.cuh); numbering = scan.cu's ScanMode
enum ScanKind : int { SCAN_EXCLUSIVE = 0, SCAN_INCLUSIVE = 1, SCAN_COMPRESS_INDEX = 2, SCAN_COMPRESS_VALUE = 3 };
constexpr int kScanFusedMaxStreams = 6;
constexpr size_t kScanFusedLagMaxNodes = 8;           // prefix sums
constexpr size_t kScanFusedLagMaxNodesCompress = 64;  // compress -> indices / values
// Geometry of a fused scan kernel (scan_fused.cuh): 1024 threads x vpt 128-bit vectors per tile.  Traces that stream at
// most one array get the lagged variant (look-back one tile behind, see scan.cu: scan_kernel_lag) unless
// $VKJIT_SCAN_IMPL=classic; the others keep the immediate look-back with a 2-slot ring per streamed array.
struct ScanFusedGeom {
  bool lag = false;
  int vpt = 6;      // 128-bit vectors per thread and tile
  int slots = 2;    // ring slots per streamed array
  int staging = 0;  // output staging tiles: lagged prefix sums 1 (TMA bulk store), lagged compress 2 (coalesced copy-out)
  int threads = 1024;  // per CTA; 512: two co-resident CTAs per SM overlap each other's per-tile chains (lagged kernels)
  int look_wide = 5;   // status words per lane and look-back round (32 x look_wide predecessors); $VKJIT_LOOK_WIDE
  // control-warp variant (compress modes, scan_fused.cuh: VK_CTRL): `threads` workers + one control warp that owns the
  // totals scan / publish / look-back chain; the workers write tile k - depth while the chain of tile k runs
  bool park = false;   // lagged prefix sums of traces that stream nothing: results wait in shared memory (VK_PARK), VPT 6
  bool ctrl = false;
  int depth = 3;       // tiles between evaluation and output (workers)
  int clag = 2;        // tiles between a tile's publish and its resolve (control warp); depth >= clag
  int ctas = 1;        // co-resident CTAs per SM the kernel is compiled for
  int launch_threads() const { return ctrl ? threads + 64 : threads; }  // + control warp + TMA producer warp
  size_t tile() const { return (size_t)threads * 4 * vpt; }
  size_t smem(size_t streams) const { return (streams * slots + (size_t)staging) * tile() * 4; }
};
ScanFusedGeom scan_fused_geom(size_t streams, int mode, size_t nodes);
constexpr int kScanWregDefault = 0;
constexpr int kScanLagPackedDefault = 0;  // packed, anchored look-back in the lagged fused scan kernels ($VKJIT_LAG_PACKED)
bool scan_lag_packed();
constexpr int kScanParkDefault = 0;  // parked-result lagged prefix sums for traces without streamed inputs ($VKJIT_SCAN_PARK)
constexpr int kScanCtrlDefault = 0;  // control-warp fused compress kernel ($VKJIT_SCAN_CTRL)
const char* fscan_trace_file();  // $VKJIT_FSCAN_TRACE (per-tile phase stamps of the lagged fused scan kernels)

struct Param {
  VarId var;     // the Binding var whose array is passed
  uint8_t use;   // ParamUse bits
  uint32_t node; // local node index
};

// One scheduled trace, canonicalised.  Reused between evals (no allocation once warm).
struct Program {
  std::vector<uint32_t> key;    // canonical words (cache key, verified on hit); only [0, key_len) is meaningful
  size_t key_len = 0;
  std::vector<VarId> order;     // post-order of var ids; index == local node id
  std::vector<Param> params;    // pointer parameters, in kernel-signature order
  std::vector<uint32_t> roots;  // local node id per scheduled var
  uint64_t n = 0;               // kernel size (record_kernel_size, internal.rs:710-729)
  bool have_n = false;
  uint64_t base = 0;            // first global lane (sharded aranges)
  bool have_base = false;
  bool sharded = false;
  bool vectorized = true;       // 128-bit ld/st variant
  int reduce = -1;              // >= 0: fused trace -> reduce kernel (VKJIT_RED_*), the single root is not stored
  int scan = -1;                // >= 0: fused trace -> scan kernel (SCAN_*): root 0 is scanned / is the compress mask
  int privatize = 0;            // variant: the first scatter_add target is (partly) privatised in shared memory: 1 per CTA, 2 split over a 2-CTA
                                // cluster, 3 bin-range passes (each launch handles the lanes whose bin lies in [bin_lo, bin_lo + kbins), all of them in shared memory)
  int sadd_param = -1;          // param index of the first scatter_add target (-1: none)
  int sadd_idx_node = -1;       // local node id of that scatter_add's index
  uint32_t n_sadd = 0, n_scatter = 0;  // side-effect nodes in the trace
  uint8_t sadd_target_use = 0;  // how else the target array is used in the trace (USE_* bits)
  bool has_gather = false;      // the trace gathers (wants L1 for its table)
  Hash128 hash;

  void clear();
};

// Walks the schedule, fills `p` (key, order, params, n) and hashes it.  Throws Error on the
// reference's panics: size mismatch (internal.rs:699-702), size-less schedule (:1202),
// gather from a non-buffer (:1054), scatter into a non-buffer (:1059-1062), struct roots.
void build_program(Ir& ir, const std::vector<VarId>& schedule, bool vectorized, Program& p, int reduce = -1,
                   int privatize = 0, int scan = -1);

// number of params the kernel streams (one word per lane)
size_t stream_count(const Program& p);

// privatised scatter_add split over a 2-CTA cluster (see program.cpp); default decided by measurement
constexpr int kSaddClusterDefault = 0;
bool sadd_cluster();

// Fingerprint of the generator build (embedded device sources + generator revision); part of the disk cache header.
uint32_t generator_fingerprint();

// CUDA C source of the kernel for `p` (entry point "vkjit_trace").
std::string generate_cuda(const Ir& ir, const Program& p);

}  // namespace vkjit
