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

// fused trace -> scan kernels (scan_fused.cuh); numbering = prims.cu's ScanMode
enum ScanKind : int { SCAN_EXCLUSIVE = 0, SCAN_INCLUSIVE = 1, SCAN_COMPRESS_INDEX = 2, SCAN_COMPRESS_VALUE = 3 };
constexpr int kScanFusedMaxStreams = 6;
// tile geometry of a fused scan: T threads x vpt 128-bit vectors; the TMA ring holds 2 stages x streams x tile
int scan_fused_threads();  // threads per CTA: 1024 (one CTA per SM); $VKJIT_SCAN_T=512 runs two 512-thread CTAs per SM (measured slower, profiles/r01_fused_scan.md)
inline int scan_fused_vpt(size_t streams) { return streams <= 1 ? 6 : (int)(6 / streams); }
inline size_t scan_fused_tile(size_t streams) { return (size_t)scan_fused_threads() * 4 * scan_fused_vpt(streams); }
inline size_t scan_fused_smem(size_t streams) { return 2 * streams * scan_fused_tile(streams) * 4; }

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
  bool privatize = false;       // variant: the first scatter_add target is partly privatised in shared memory
  int sadd_param = -1;          // param index of the first scatter_add target (-1: none)
  bool has_gather = false;      // the trace gathers (wants L1 for its table)
  Hash128 hash;

  void clear();
};

// Walks the schedule, fills `p` (key, order, params, n) and hashes it.  Throws Error on the
// reference's panics: size mismatch (internal.rs:699-702), size-less schedule (:1202),
// gather from a non-buffer (:1054), scatter into a non-buffer (:1059-1062), struct roots.
void build_program(Ir& ir, const std::vector<VarId>& schedule, bool vectorized, Program& p, int reduce = -1,
                   bool privatize = false, int scan = -1);

// number of params the kernel streams (one word per lane)
size_t stream_count(const Program& p);

// CUDA C source of the kernel for `p` (entry point "vkjit_trace").
std::string generate_cuda(const Ir& ir, const Program& p);

}  // namespace vkjit
