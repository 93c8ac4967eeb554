// program.cpp — trace walk, canonical key + hash, CUDA C generation.
// Reference counterparts: Kernel::record_kernel_size (internal.rs:710-729), record_ops
// (:874-1117), compile (:1148-1305).  See program.h.
#include "program.h"
#include <string>

#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <cmath>

namespace vkjit {

void Program::clear() {
  key_len = 0; order.clear(); params.clear(); roots.clear();
  n = 0; have_n = false; base = 0; have_base = false; sharded = false; vectorized = true; reduce = -1; scan = -1;
  privatize = 0; sadd_param = -1; has_gather = false; sadd_idx_node = -1; n_sadd = 0; n_scatter = 0; sadd_target_use = 0;
  hash = Hash128();
}

namespace {

struct Frame { VarId id; uint32_t next; };

// internal.rs:697-706
inline void set_num(Program& p, uint64_t n) {
  if (p.have_n) {
    if (p.n != n)
      fail(VKJIT_ERR_SIZE, "All variables in the kernel have to have the same number of elements! (" +
                               std::to_string(p.n) + " vs " + std::to_string(n) + ", internal.rs:699-702)");
  } else { p.have_n = true; p.n = n; }
}

// Child k of a node in the order record_ops recurses (internal.rs:874-1117); returns false
// when there is no k-th child.  `ptr_only` marks operands that are addressed, not evaluated.
inline bool child_of(const Var& v, uint32_t k, VarId& out, uint8_t& ptr_use) {
  ptr_use = 0;
  const VarId* d = v.deps();
  switch (v.op) {
    case OP_GATHER:
      if (k == 0) { out = d[0]; ptr_use = USE_GATHER; return true; }
      if (k < v.ndeps) { out = d[k]; return true; }
      return false;
    case OP_SCATTER: case OP_SCATTER_ADD:
      if (k == 0) { out = d[0]; return true; }
      if (k == 1) { out = v.side_effect; ptr_use = USE_SCATTER; return true; }
      if (k - 1 < v.ndeps) { out = d[k - 1]; return true; }
      return false;
    default:
      if (k < v.ndeps) { out = d[k]; return true; }
      return false;
  }
}

void type_signature(const Ir& ir, TypeId t, std::vector<uint32_t>& key) {
  if (!ty_is_struct(t)) { key.push_back(t); return; }
  const auto& e = ir.struct_elems(t);
  key.push_back(0x80000000u | (uint32_t)e.size());
  for (TypeId x : e) type_signature(ir, x, key);
}

constexpr size_t kMaxVectorNodes = 96;

// $VKJIT_FAST_MATH=1: exp/log/sin/cos lower to CUDA's expf/logf/sinf/cosf (MUFU-assisted, 1-2 ulp, NOT bit-identical to
// the oracle) instead of vk_math.h.  A speed knob for users who do not need reproducible bits; never the default,
// never what the parity tests run.  Part of the cache key.
int g_fast_math = -1;
bool fast_math() {
  if (g_fast_math < 0) { const char* s = getenv("VKJIT_FAST_MATH"); g_fast_math = (s && s[0] == '1') ? 1 : 0; }
  return g_fast_math == 1;
}

// $VKJIT_NO_RANGES=1: never use the range-proven unchecked vk_math.h fast paths (A/B knob).  Part of the cache key.
int g_no_ranges = -1;
bool no_ranges() {
  if (g_no_ranges < 0) { const char* s = getenv("VKJIT_NO_RANGES"); g_no_ranges = (s && s[0] == '1') ? 1 : 0; }
  return g_no_ranges == 1;
}

// Warp aggregation of integer scatter_add (kSaddHelper) is OFF unless $VKJIT_AGG=1: measured on B200
// (profiles/r02_h26_hot_bins.md) it loses everywhere — shared-memory privatisation already absorbs hot bins (2^26 lanes
// into 1..4096 bins: 0.096-0.100 ms with plain ATOMS, ptxas folds warp-uniform addresses into one REDUX + ATOMS itself),
// while MATCH.ANY + REDUX per scatter_add cost 0.11 / 0.65 / 0.48 ms for 1 / 16 / 256 bins.  Part of the cache key.
int g_no_agg = -1;
bool no_agg() {
  if (g_no_agg < 0) { const char* s = getenv("VKJIT_AGG"); g_no_agg = (s && s[0] == '1') ? 0 : 1; }
  return g_no_agg == 1;
}

}  // namespace
// Privatised scatter_add across a 2-CTA thread-block cluster: each CTA keeps HALF of the privatised bins in its own
// shared memory and reaches the other half through distributed shared memory (mapa + red.shared::cluster), so a
// 2^16-bin target (256 KiB) is privatised completely (no L2 RED at all) and each SM keeps ~100 KB of L1 for gathers.
// $VKJIT_SADD_CLUSTER=0/1 overrides the default; part of the cache key.
bool sadd_cluster() {
  static const int on = [] { const char* s = getenv("VKJIT_SADD_CLUSTER"); return s ? (s[0] == '1' ? 1 : 0) : kSaddClusterDefault; }();
  return on == 1;
}
namespace {

int g_unroll = -1;
int unroll_factor() {
  if (g_unroll < 0) {
    const char* s = getenv("VKJIT_UNROLL");
    g_unroll = s ? std::max(1, atoi(s)) : 2;
  }
  return g_unroll;
}

}  // namespace

ScanFusedGeom scan_fused_geom(size_t streams, int mode, size_t nodes) {
  static int classic = -1;
  if (classic < 0) { const char* s = getenv("VKJIT_SCAN_IMPL"); classic = (s && std::string(s) == "classic") ? 1 : 0; }
  ScanFusedGeom g;
  // Lagging pays when the tile time is dominated by the look-back latency.  Prefix sums of ALU-heavy bodies (a hash
  // per lane) are issue bound and do better with the larger tiles of the immediate look-back (profiles/r01_fused_scan.md);
  // the compress modes take the lagged kernel up to 64 nodes since it runs as TWO 512-thread CTAs per SM
  // (profiles/r02_fused_scan.md: the per-tile chain evaluate -> scan -> barrier -> resolve -> barrier -> stores of one
  // CTA leaves the SM idle 57 % of the time; a second co-resident CTA fills it: compress_values(v, v > t) 0.436 -> 0.378 ms,
  // compress_values(v, hash mask) 0.466 -> 0.415 ms).  Prefix sums got SLOWER that way (0.361 -> 0.507 ms: the staging
  // tile + TMA bulk store want the big tile) and keep one 1024-thread CTA.
  static size_t lag_max_nodes = 0;
  if (!lag_max_nodes) { const char* e = getenv("VKJIT_LAG_MAX_NODES"); lag_max_nodes = e ? (size_t)std::max(1, atoi(e)) : 0; }
  const size_t max_nodes = lag_max_nodes ? lag_max_nodes : (mode >= SCAN_COMPRESS_INDEX ? kScanFusedLagMaxNodesCompress : kScanFusedLagMaxNodes);
  if (streams <= 1 && nodes <= max_nodes && !classic) {
    g.lag = true;
    if (mode >= SCAN_COMPRESS_INDEX) { g.threads = 512; g.vpt = 4; g.slots = 3; }  // values: slot k-1 kept, k current, k+1 in flight
    else { g.vpt = 4; g.slots = 2; g.staging = 1; }
    // experiments / tuning: $VKJIT_SCAN_T (512: two CTAs per SM), $VKJIT_SCAN_VPT, $VKJIT_SCAN_SLOTS
    static int t_env = -1, vpt_env = -1, slots_env = -1;
    if (t_env < 0) {
      const char* e = getenv("VKJIT_SCAN_T"); t_env = e ? atoi(e) : 0;
      e = getenv("VKJIT_SCAN_VPT"); vpt_env = e ? atoi(e) : 0;
      e = getenv("VKJIT_SCAN_SLOTS"); slots_env = e ? atoi(e) : 0;
    }
    if (t_env == 512 || t_env == 1024) g.threads = t_env;
    if (vpt_env > 0) g.vpt = vpt_env;
    if (slots_env > 0) g.slots = slots_env;
  } else {
    g.vpt = streams <= 1 ? 6 : (int)(6 / streams);
    static int t_nolag = -1;
    if (t_nolag < 0) { const char* e = getenv("VKJIT_SCAN_T_NOLAG"); t_nolag = e ? atoi(e) : 0; }
    if (t_nolag == 512 || t_nolag == 1024) g.threads = t_nolag;
  }
  // Prefix sums of traces that stream NOTHING (e.g. prefix_sum(hash(lane))): no input ring, so two tile-sized buffers
  // of shared memory park the row-relative results while the look-back lags one tile behind — large tiles (VPT 6, as the
  // immediate-look-back kernel) AND a lagged look-back.  $VKJIT_SCAN_PARK=0/1.
  {
    static const int park_env = [] { const char* e = getenv("VKJIT_SCAN_PARK"); return e ? (e[0] == '1' ? 1 : 0) : kScanParkDefault; }();
    if (park_env == 1 && mode < SCAN_COMPRESS_INDEX && streams == 0 && nodes <= kScanFusedLagMaxNodesCompress && !classic) {
      g = ScanFusedGeom();
      g.lag = true; g.park = true; g.threads = 1024; g.vpt = 6; g.slots = 2; g.staging = 2;  // (slots only sizes arrays: no streams)
      static const int pt = [] { const char* e = getenv("VKJIT_PARK_T"); return e ? atoi(e) : 0; }();
      static const int pv = [] { const char* e = getenv("VKJIT_PARK_VPT"); return e ? atoi(e) : 0; }();
      if (pt == 512 || pt == 1024) g.threads = pt;
      if (pv >= 1 && pv <= 8 && (g.threads / 32 * pv) % 32 == 0) g.vpt = pv;
    }
  }
  // $VKJIT_SCAN_CTRL=1 (experiment): compress modes over at most one streamed array take the control-warp kernel.
  // Geometry through $VKJIT_CTRL_T (workers), $VKJIT_CTRL_VPT, $VKJIT_CTRL_SLOTS, $VKJIT_CTRL_DEPTH, $VKJIT_CTRL_CTAS.
  {
    static int ctrl_env = -1, ct = 0, cv = 0, cs = 0, cd = 0, cc = 0, cl = 0;
    if (ctrl_env < 0) {
      const char* e = getenv("VKJIT_SCAN_CTRL"); ctrl_env = e ? (e[0] == '1' ? 1 : 0) : kScanCtrlDefault;
      e = getenv("VKJIT_CTRL_T"); ct = e ? atoi(e) : 0;
      e = getenv("VKJIT_CTRL_VPT"); cv = e ? atoi(e) : 0;
      e = getenv("VKJIT_CTRL_SLOTS"); cs = e ? atoi(e) : 0;
      e = getenv("VKJIT_CTRL_DEPTH"); cd = e ? atoi(e) : 0;
      e = getenv("VKJIT_CTRL_CTAS"); cc = e ? atoi(e) : 0;
      e = getenv("VKJIT_CTRL_LAG"); cl = e ? atoi(e) : 0;
    }
    if (ctrl_env == 1 && mode >= SCAN_COMPRESS_INDEX && streams <= 1 && nodes <= max_nodes && !classic) {
      const bool values = mode == SCAN_COMPRESS_VALUE;
      g.ctrl = true; g.lag = true; g.staging = 0;
      g.threads = (ct == 256 || ct == 512 || ct == 1024) ? ct : 512;
      g.vpt = (cv >= 1 && cv <= 4) ? cv : 4;
      g.clag = (cl >= 1 && cl <= 12) ? cl : 2;
      g.depth = (cd >= 1 && cd <= 13) ? cd : g.clag + 1;
      if (g.depth < g.clag) g.depth = g.clag;
      g.slots = cs > 0 ? cs : (values ? g.depth + 2 : 3);
      if (values && g.slots < g.depth + 2) g.slots = g.depth + 2;
      if (g.slots < 2) g.slots = 2;
      // the ring must fit next to ~14 KB of static shared memory: smaller tiles first, then less depth
      while (streams && (size_t)g.slots * g.tile() * 4 > 208 * 1024) {
        if (g.vpt > 2 && g.threads * (g.vpt / 2) % 1024 == 0) g.vpt /= 2;
        else if (values && g.depth > g.clag) { g.depth -= 1; g.slots = g.depth + 2; }
        else if (values && g.clag > 1) { g.clag -= 1; g.depth = g.clag; g.slots = g.depth + 2; }
        else if (g.slots > 2) g.slots -= 1;
        else break;
      }
      // co-resident CTAs: what 227 KB of shared memory and 2048 threads allow, at most 2
      const size_t smem = std::max<size_t>(streams, 0) * g.slots * g.tile() * 4 + 8192;
      int fit = (int)std::min<size_t>(2, std::min<size_t>((227 * 1024) / std::max<size_t>(smem, 1), 2048 / (size_t)(g.threads + 64)));
      if (fit < 1) fit = 1;
      g.ctas = (cc == 1 || cc == 2) ? std::min(cc, fit) : fit;
    }
  }
  // 160 predecessors per look-back round.  Two 512-thread CTAs per SM make 296 tiles per generation, so the upper half of
  // a generation needs a second round — yet a 320-wide window measured SLOWER (profiles/r02_fused_scan.md, experiment 3:
  // compress_values(v, v > t) 0.376 -> 0.410 ms, 384-wide 0.469 ms): the status traffic costs more than the second round.
  g.look_wide = 5;  // (control-warp kernels read a packed window of up to 318 aggregates through the same 5 x 32 x 16 bytes)
  static int lw_env = -1;
  if (lw_env < 0) { const char* e = getenv("VKJIT_LOOK_WIDE"); lw_env = e ? atoi(e) : 0; }
  if (lw_env >= 1 && lw_env <= 16) g.look_wide = lw_env;
  return g;
}

// $VKJIT_FSCAN_TRACE=<file>: the lagged fused scan kernels stamp %globaltimer per tile and phase (scan_fused.cuh: VK_STAMP);
// runtime.cpp: eval_scan writes the stamps to the file after every launch.  Part of the cache key.
const char* fscan_trace_file() {
  static const char* f = getenv("VKJIT_FSCAN_TRACE");
  return (f && f[0]) ? f : nullptr;
}

// $VKJIT_SCAN_EARLY=1: the lagged fused scan kernels request a tile's status window one phase EARLIER than at the start
// of the iteration that resolves it.  Measured slower (profiles/r02_fused_scan.md, experiment 4: the predecessors of the
// same generation have not published yet, the early window is full of INVALID words that must be polled again).
// A/B knob, default off, part of the cache key.
bool scan_early() {
  static const int on = [] { const char* e = getenv("VKJIT_SCAN_EARLY"); return (e && e[0] == '1') ? 1 : 0; }();
  return on == 1;
}

// $VKJIT_SCAN_WREG=0/1: the lagged fused scan kernels read a tile's status window with strong loads into registers
// (scan_fused.cuh: VK_WREG) instead of cp.async.cg into shared memory.  Part of the cache key.
bool scan_wreg() {
  static const int on = [] { const char* e = getenv("VKJIT_SCAN_WREG"); return e ? (e[0] == '1' ? 1 : 0) : kScanWregDefault; }();
  return on == 1;
}

// $VKJIT_FSCAN_DIAG bit 0: timing-only diagnostic of the control-warp compress kernels (no output stores; WRONG results);
// bit 1: staged, coalesced warp-row output instead of lane-by-lane stores (A/B; measured slower).  Part of the key.
int scan_diag() {
  static const int d = [] { const char* e = getenv("VKJIT_FSCAN_DIAG"); return e ? atoi(e) : 0; }();
  return d;
}

// $VKJIT_LAG_PACKED=0/1: the lagged fused scan kernels resolve a tile's prefix from a packed copy of the tile aggregates,
// anchored at the CTA's own previous tile (scan_fused.cuh: VK_LAGPACK).  Part of the cache key.
bool scan_lag_packed() {
  static const int on = [] { const char* e = getenv("VKJIT_LAG_PACKED"); return e ? (e[0] == '1' ? 1 : 0) : kScanLagPackedDefault; }();
  return on == 1;
}

size_t stream_count(const Program& p) {
  size_t k = 0;
  for (const Param& pr : p.params) k += (pr.use & USE_STREAM) ? 1 : 0;
  return k;
}

void build_program(Ir& ir, const std::vector<VarId>& schedule, bool vectorized, Program& p, int reduce, int privatize, int scan) {
  p.clear();
  p.vectorized = vectorized;
  p.reduce = reduce;
  p.scan = scan;
  p.privatize = privatize;
  const uint32_t stamp = ir.next_stamp();
  // -fPIC: every access to a thread_local goes through __tls_get_addr, so the walk binds them once
  // (measured on the 364-node Monte-Carlo trace: 17.5 -> 12 us)
  static thread_local std::vector<Frame> tls_stack;
  static thread_local std::vector<uint32_t> tls_sig, tls_binding_pos;
  std::vector<Frame>& stackv = tls_stack;
  std::vector<uint32_t>& sig = tls_sig;
  std::vector<uint32_t>& binding_pos = tls_binding_pos;  // key index of each param's word (use bits patched at the end)
  binding_pos.clear();
  if (stackv.size() < 256) stackv.resize(256);
  Frame* stk = stackv.data();  // explicit DFS stack with a raw cursor
  size_t sp = 0, scap = stackv.size();
  Var* const vars = ir.vars.data();
  const size_t nvars = ir.vars.size();

  // canonical key: structure only — no VarIds, no addresses, no n.  Words are emitted
  // while nodes are numbered (single pass; this walk is the cache-hit critical path).
  std::vector<uint32_t>& key = p.key;
  // raw write cursor into the key buffer: push_back through the Program reference reloads and stores
  // the vector's end pointer on every word (measured 36 -> 20 ns per node)
  size_t kn = 2;
  if (key.size() < 4096) key.resize(4096);  // the buffer keeps its size between walks; p.key_len is the logical length
  uint32_t* kw = key.data();
  kw[0] = 0x564B4A32u;  // "VKJ2"; kw[1] = variant word, patched below
  std::vector<VarId>& order = p.order;
  // node word: op | kind << 8 | min(ndeps, 255) << 16 | type << 24 | has_side_effect << 28
  auto number_node = [&](VarId id, Var& v) {
    v.stamp = stamp; v.local = (uint32_t)order.size();
    order.push_back(id);
    const uint32_t nd = v.ndeps;
    const uint32_t tycode = ty_is_struct(v.ty) ? 0xFu : v.ty;
    if (kn + 16 + nd > key.size()) { key.resize(std::max(key.size() * 2, kn + 64 + nd)); kw = key.data(); }
    kw[kn++] = (uint32_t)v.op | ((uint32_t)(v.kind & 0xFFu) << 8) | ((nd < 255u ? nd : 255u) << 16) | (tycode << 24) |
               ((uint32_t)(v.has_se ? 1u : 0u) << 28);
    if (nd >= 255u) kw[kn++] = nd;
    if (tycode == 0xFu) {  // struct-typed node: variable-length type signature
      sig.clear();
      type_signature(ir, v.ty, sig);
      if (kn + 16 + nd + sig.size() > key.size()) { key.resize(std::max(key.size() * 2, kn + 64 + nd + sig.size())); kw = key.data(); }
      for (uint32_t w : sig) kw[kn++] = w;
    }
    switch (v.op) {
      case OP_CONST: case OP_GETATTR: case OP_SETATTR: kw[kn++] = v.aux; break;
      case OP_ARANGE: kw[kn++] = v.sharded; break;
      case OP_BINDING:
        v.aux = (uint32_t)p.params.size();
        p.params.push_back({id, 0, v.local});
        binding_pos.push_back((uint32_t)kn);
        kw[kn++] = v.aux;
        break;
      default: break;
    }
    const VarId* d = v.deps();
    for (uint32_t k = 0; k < nd; ++k) kw[kn++] = vars[d[k]].local;
    if (v.has_se) kw[kn++] = vars[v.side_effect].local;
  };

  auto touch_binding = [&](Var& v, uint8_t use) {
    Param& pr = p.params[v.aux];
    if ((use & USE_STREAM) && !(pr.use & USE_STREAM)) {  // internal.rs:717-721
      set_num(p, v.array->bytes / 4);
      if (v.sharded) p.sharded = true;
    }
    pr.use |= use;
  };
  auto number_arange = [&](VarId id, Var& v) {  // internal.rs:722-725
    number_node(id, v);
    set_num(p, v.num);
    if (v.sharded) {
      p.sharded = true;
      if (p.have_base && p.base != v.base) fail(VKJIT_ERR_SIZE, "sharded aranges with different bases in one kernel");
      p.have_base = true; p.base = v.base;
    }
  };

  for (VarId root : schedule) {
    {
      const Var& rv = ir.var(root);
      if (!ty_is_scalar(rv.ty))  // Struct: stride() is unimplemented!() (vartype.rs:45-53)
        fail(VKJIT_ERR_UNSUPPORTED, "only scalar-typed vars can be scheduled (vartype.rs:45-53)");
    }
    if (vars[root].stamp == stamp) {
      if (vars[root].op == OP_BINDING) touch_binding(vars[root], USE_STREAM);
      p.roots.push_back(vars[root].local);
      continue;
    }
    stk[sp++] = {root, 0};  // sp == 0 here and scap >= 256
    while (sp) {
    next_frame:
      Frame& f = stk[sp - 1];
      Var& v = vars[f.id];
      const bool indexed = v.op == OP_GATHER || v.op == OP_SCATTER || v.op == OP_SCATTER_ADD;
      for (;;) {
        VarId c; uint8_t ptr_use = 0;
        if (!indexed) {  // fast path: children are the deps in order
          if (f.next >= v.ndeps) break;
          c = v.deps()[f.next++];
        } else {
          if (!child_of(v, f.next, c, ptr_use)) break;
          ++f.next;
        }
        if (c >= nvars || vars[c].op == OP_FREE) fail(VKJIT_ERR_INVALID, "invalid VarId " + std::to_string(c));
        Var& cv = vars[c];
        if (ptr_use) {
          if (ptr_use == USE_GATHER && (cv.op != OP_BINDING || !cv.array))
            fail(VKJIT_ERR_INVALID, "Can only gather from buffer! (internal.rs:1054)");
          if (ptr_use == USE_SCATTER && (cv.op != OP_BINDING || !cv.array))
            fail(VKJIT_ERR_INVALID, "Cannot scatter into non buffer variables! (internal.rs:1061)");
          if (cv.stamp != stamp) number_node(c, cv);
          touch_binding(cv, ptr_use);
        } else if (cv.stamp == stamp) {
          if (cv.op == OP_BINDING) touch_binding(cv, USE_STREAM);
        } else if (cv.ndeps == 0 && !cv.has_se) {  // leaf: number in place, no frame
          if (cv.op == OP_BINDING) {
            if (!cv.array) fail(VKJIT_ERR_INVALID, "Binding without an array");
            number_node(c, cv);
            touch_binding(cv, USE_STREAM);
          } else if (cv.op == OP_ARANGE) {
            number_arange(c, cv);
          } else {
            number_node(c, cv);
          }
        } else {
          if (sp == scap) { stackv.resize(scap * 2); stk = stackv.data(); scap = stackv.size(); }
          stk[sp++] = {c, 0};  // f is stale from here on
          goto next_frame;
        }
      }
      // all children done: number this node (post-order)
      const VarId id = f.id;
      --sp;
      if (v.stamp == stamp) continue;
      switch (v.op) {
        case OP_BINDING:
          if (!v.array) fail(VKJIT_ERR_INVALID, "Binding without an array");
          number_node(id, v);
          touch_binding(v, USE_STREAM);
          break;
        case OP_ARANGE: number_arange(id, v); break;
        case OP_GATHER: case OP_SCATTER: case OP_SCATTER_ADD: {
          const TypeId it = vars[v.deps()[1]].ty;
          if (it != VKJIT_TY_U32 && it != VKJIT_TY_I32) fail(VKJIT_ERR_TYPE, "gather/scatter index must be U32 or I32");
          if (v.ndeps >= 3 && vars[v.deps()[2]].ty != VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "gather/scatter mask must be Bool");
          if (!ty_is_scalar(v.ty)) fail(VKJIT_ERR_UNSUPPORTED, "gather/scatter of a struct");
          if (v.op != OP_GATHER && vars[v.side_effect].ty != v.ty) fail(VKJIT_ERR_TYPE, "scatter: source and target types differ");
          if (v.op == OP_SCATTER_ADD && v.ty == VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "scatter_add on Bool");
          if (v.op == OP_SCATTER_ADD && p.sadd_param < 0) { p.sadd_param = (int)vars[v.side_effect].aux; p.sadd_idx_node = (int)vars[v.deps()[1]].local; }
          if (v.op == OP_SCATTER_ADD) p.n_sadd += 1;
          if (v.op == OP_SCATTER) p.n_scatter += 1;
          if (v.op == OP_GATHER) p.has_gather = true;
          number_node(id, v);
          break;
        }
        case OP_SELECT:
          if (vars[v.deps()[0]].ty != VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "select condition must be Bool");
          number_node(id, v);
          break;
        default: number_node(id, v); break;
      }
    }
    p.roots.push_back(vars[root].local);
  }

  // ALU-heavy traces (config 5: ~200 ops/lane) are bound by SM issue rate, not by HBM: 128-bit
  // accesses buy nothing there, while inlining the body 4x (+ unrolling) multiplies NVRTC time
  // and I-cache footprint (measured: 3.1 s / 789 KB cubin vs 0.4 s).  They get the one-copy variant.
  if (p.order.size() > kMaxVectorNodes) p.vectorized = false;
  vectorized = p.vectorized;

  if (!p.have_n) fail(VKJIT_ERR_SIZE, "schedule has no Binding/Arange: kernel size unknown (internal.rs:1202 num.unwrap())");
  if (p.n > 0xFFFFFFFFull) fail(VKJIT_ERR_SIZE, "kernel size exceeds the 32-bit invocation index");
  if (p.params.size() + p.roots.size() > 480) fail(VKJIT_ERR_UNSUPPORTED, "too many arrays in one kernel (4 KB parameter limit)");

  if (kn + 12 + p.roots.size() > key.size()) { key.resize(kn + 64 + p.roots.size()); kw = key.data(); }
  if (p.sadd_param < 0 || reduce >= 0 || scan >= 0) p.privatize = 0;
  if (p.sadd_param >= 0) p.sadd_target_use = p.params[p.sadd_param].use;
  if (p.privatize == 3) {
    // bin-range passes predicate the WHOLE lane body on the bin: only sound when that scatter_add is the trace's one
    // side effect and its target is not read in the same trace
    if (p.n_sadd != 1 || p.n_scatter != 0 || (p.sadd_target_use & (USE_STREAM | USE_GATHER))) p.privatize = 1;
  }
  kw[1] = (vectorized ? 1u : 0u) | (no_ranges() ? 2u : 0u) | ((uint32_t)unroll_factor() << 8) | ((uint32_t)(reduce + 1) << 16) | (p.privatize ? 1u << 24 : 0u) |
          ((uint32_t)(scan + 1) << 25) | (fast_math() ? 1u << 29 : 0u) | (no_agg() ? 1u << 30 : 0u) | (p.privatize == 2 ? 1u << 31 : 0u) | (p.privatize == 3 ? 4u : 0u);  // bit 28 (lagged fused scan) is patched once the streams are known
  for (size_t k = 0; k < p.params.size(); ++k) kw[binding_pos[k]] |= (uint32_t)p.params[k].use << 16;
  if (scan >= 0) {
    const ScanFusedGeom sg = scan_fused_geom(stream_count(p), scan, p.order.size());
    if (sg.lag) kw[1] |= 1u << 28;
    kw[kn++] = 0xFFFFFFFEu;  // geometry of the fused scan kernel (tunable through the environment)
    kw[kn++] = (uint32_t)sg.threads | ((uint32_t)sg.look_wide << 11) | ((uint32_t)sg.vpt << 16) | ((uint32_t)sg.slots << 24);
    kw[kn++] = (sg.ctrl ? 1u : 0u) | (sg.park ? 2u : 0u) | ((sg.lag && !sg.ctrl && scan_lag_packed()) ? 4u : 0u) | ((uint32_t)sg.depth << 4) | ((uint32_t)sg.ctas << 8) | ((uint32_t)sg.clag << 12) |
               (fscan_trace_file() ? 1u << 30 : 0u) | (scan_early() ? 1u << 29 : 0u) | (scan_wreg() ? 1u << 28 : 0u) | ((uint32_t)(scan_diag() & 3) << 26) | ((uint32_t)((scan_diag() >> 2) & 1) << 25);
  }
  kw[kn++] = 0xFFFFFFFFu;
  for (uint32_t r : p.roots) kw[kn++] = r;
  while (kn & 7) kw[kn++] = 0u;
  p.key_len = kn;
  // 128-bit hash: four independent multiply chains over 64-bit pairs (a single chain is latency bound: ~8 cycles
  // per pair, 2.5 us for the 364-node trace), folded at the end.  The full key is compared on every cache hit anyway.
  uint64_t a0 = 0x243F6A8885A308D3ull, a1 = 0x13198A2E03707344ull, a2 = 0xA4093822299F31D0ull, a3 = 0x082EFA98EC4E6C89ull;
  const uint32_t* kr = key.data();
  for (size_t i = 0; i < kn; i += 8) {
    uint64_t w[4];
    memcpy(w, kr + i, 32);  // four 64-bit pairs (memcpy: the words were written through a uint32_t pointer)
    a0 = (a0 ^ w[0]) * 0x9E3779B97F4A7C15ull; a0 ^= a0 >> 32;
    a1 = (a1 ^ w[1]) * 0xC2B2AE3D27D4EB4Full; a1 ^= a1 >> 29;
    a2 = (a2 ^ w[2]) * 0x165667B19E3779F9ull; a2 ^= a2 >> 31;
    a3 = (a3 ^ w[3]) * 0xD6E8FEB86659FD93ull; a3 ^= a3 >> 30;
  }
  uint64_t h0 = (a0 ^ (a1 * 0x9E3779B97F4A7C15ull)) * 0xC2B2AE3D27D4EB4Full; h0 ^= h0 >> 29;
  h0 = (h0 ^ a2) * 0x165667B19E3779F9ull; h0 ^= h0 >> 32;
  uint64_t h1 = (a3 + (a2 * 0xD6E8FEB86659FD93ull)) * 0x9E3779B97F4A7C15ull; h1 ^= h1 >> 31;
  h1 = (h1 ^ a0 ^ (a1 >> 7)) * 0xC2B2AE3D27D4EB4Full; h1 ^= h1 >> 29;
  p.hash.lo = h0 ^ kn;
  p.hash.hi = h1;
}

// ---------------------------------------------------------------------------------------
// CUDA C generation
// ---------------------------------------------------------------------------------------
namespace {

// Value ranges the generator can PROVE from the trace alone (constants, lane indices, shifts, casts, f32 arithmetic
// through monotone rounding, the documented <= 1 ulp bounds of vk_math.h).  Used for one thing: calling the unchecked
// fast path of a vk_math.h function when its argument provably never takes the out-of-line path (NaN / inf / huge /
// subnormal / -0 handling) — same bits, no test, no branch, and the compiler can schedule across the call.  Nothing here
// depends on n, addresses or data, so a cached kernel stays valid for every launch of its canonical key.
struct FRange {            // f32: every value is finite, not NaN and inside [lo, hi]
  bool ok = false;
  float lo = 0.0f, hi = 0.0f;
  bool negzero = true;     // the value may be -0.0 (sin(-0) = -0 is the one thing the trig fast path gets wrong)
};
struct URange { uint32_t lo = 0u, hi = 0xFFFFFFFFu; };   // u32: always valid

struct Val {
  std::string name;        // scalar expression name
  TypeId ty = VKJIT_TY_VOID;
  std::vector<Val> elems;  // struct members (scalar replacement: structs never reach memory)
  FRange fr;
  URange ur;
};

inline float f_down(float x, int steps) { for (int i = 0; i < steps; ++i) x = std::nextafterf(x, -INFINITY); return x; }
inline float f_up(float x, int steps) { for (int i = 0; i < steps; ++i) x = std::nextafterf(x, INFINITY); return x; }
inline bool f_finite(float x) { return std::isfinite(x); }
inline FRange fr_make(float lo, float hi, bool negzero) {
  FRange r;
  r.ok = f_finite(lo) && f_finite(hi) && lo <= hi;
  r.lo = lo; r.hi = hi; r.negzero = negzero;
  return r;
}

const char* ctype(TypeId t) {
  switch (t) {
    case VKJIT_TY_BOOL: return "bool";
    case VKJIT_TY_U32: return "u32";
    case VKJIT_TY_I32: return "i32";
    case VKJIT_TY_F32: return "f32";
    default: return "void";
  }
}

std::string hex32(uint32_t w) { char b[16]; snprintf(b, sizeof b, "0x%08xu", w); return b; }

// 4-byte word -> typed value and back (Bool arrays hold 0/1 words, vartype.rs:45-64)
std::string from_word(TypeId t, const std::string& w) {
  switch (t) {
    case VKJIT_TY_BOOL: return "(" + w + " != 0u)";
    case VKJIT_TY_I32: return "(i32)" + w;
    case VKJIT_TY_F32: return "__uint_as_float(" + w + ")";
    default: return w;
  }
}
std::string to_word(TypeId t, const std::string& v) {
  switch (t) {
    case VKJIT_TY_BOOL: return "(" + v + " ? 1u : 0u)";
    case VKJIT_TY_I32: return "(u32)" + v;
    case VKJIT_TY_F32: return "__float_as_uint(" + v + ")";
    default: return v;
  }
}

struct Gen {
  const Ir& ir;
  const Program& p;
  std::string body;
  std::vector<Val> vals;

  std::vector<int> trig_partner;  // sin(x)/cos(x) of the same x: local id of the sibling, else -1
  bool uses_vk_math = false;      // the kernel text needs vk_math.h
  bool uses_agg = false;          // integer scatter_add: vk_lane carries the warp-aggregation state

  Gen(const Ir& i, const Program& pr) : ir(i), p(pr) {
    vals.resize(pr.order.size());
    // Box-Muller style traces take sin and cos of the same value: pair them so that one sincosf
    // (one range reduction) serves both
    trig_partner.assign(pr.order.size(), -1);
    std::vector<int> first_sin(pr.order.size(), -1), first_cos(pr.order.size(), -1);
    for (uint32_t li = 0; li < pr.order.size(); ++li) {
      const Var& v = ir.vars[pr.order[li]];
      if (v.op != OP_UOP || (v.kind != VKJIT_UOP_SIN && v.kind != VKJIT_UOP_COS)) continue;
      const uint32_t d = ir.vars[v.deps()[0]].local;
      std::vector<int>& mine = v.kind == VKJIT_UOP_SIN ? first_sin : first_cos;
      std::vector<int>& other = v.kind == VKJIT_UOP_SIN ? first_cos : first_sin;
      if (mine[d] < 0) mine[d] = (int)li;
      if (other[d] >= 0 && trig_partner[other[d]] < 0 && trig_partner[li] < 0 && mine[d] == (int)li) {
        trig_partner[li] = other[d];
        trig_partner[other[d]] = (int)li;
      }
    }
  }

  void line(const std::string& s) { body += "  " + s + "\n"; }
  void def(uint32_t li, TypeId ty, const std::string& expr) {
    vals[li].name = "v" + std::to_string(li);
    vals[li].ty = ty;
    line(std::string("const ") + ctype(ty) + " v" + std::to_string(li) + " = " + expr + ";");
  }
  const Val& dep(const Var& v, uint32_t k) const { return vals[ir.vars[v.deps()[k]].local]; }

  // componentwise select for struct values (reference: OpSelect-by-branch, internal.rs:1096-1112)
  Val select_val(const std::string& c, const Val& a, const Val& b, const std::string& name, int& counter) {
    Val r; r.ty = a.ty;
    if (a.elems.empty() && ty_is_scalar(a.ty)) {
      r.name = name + (counter ? "_" + std::to_string(counter) : "");
      ++counter;
      line(std::string("const ") + ctype(a.ty) + " " + r.name + " = " + c + " ? " + a.name + " : " + b.name + ";");
    } else {
      for (size_t i = 0; i < a.elems.size(); ++i) r.elems.push_back(select_val(c, a.elems[i], b.elems[i], name, counter));
    }
    return r;
  }

  std::string bop_expr(const Var& v, const std::string& a, const std::string& b) {
    const TypeId ot = ir.vars[v.deps()[0]].ty;  // operand type after promotion (internal.rs:887, :914)
    const bool f = ot == VKJIT_TY_F32, s = ot == VKJIT_TY_I32, bl = ot == VKJIT_TY_BOOL;
    auto wrap = [&](const char* op) {  // modular 2^32 arithmetic; signed overflow is UB in C
      return s ? "(i32)((u32)" + a + " " + op + " (u32)" + b + ")" : a + " " + op + " " + b;
    };
    switch (v.kind) {
      // f32: round-to-nearest intrinsics are never contracted into FMA, whatever the flags
      case VKJIT_BOP_ADD: return f ? "__fadd_rn(" + a + ", " + b + ")" : wrap("+");
      case VKJIT_BOP_SUB: return f ? "__fsub_rn(" + a + ", " + b + ")" : wrap("-");
      case VKJIT_BOP_MUL: return f ? "__fmul_rn(" + a + ", " + b + ")" : wrap("*");
      case VKJIT_BOP_DIV: return f ? "__fdiv_rn(" + a + ", " + b + ")" : a + " / " + b;  // OpFDiv / OpSDiv / OpUDiv
      case VKJIT_BOP_LT: return a + " < " + b;     // FOrdLessThan / SLessThan / ULessThan by operand type
      case VKJIT_BOP_GT: return a + " > " + b;
      case VKJIT_BOP_LEQ: return a + " <= " + b;
      case VKJIT_BOP_GEQ: return a + " >= " + b;
      case VKJIT_BOP_EQ: return a + " == " + b;    // FOrdEqual: false on NaN, as C ==
      case VKJIT_BOP_NEQ: return f ? "(" + a + " < " + b + ") || (" + a + " > " + b + ")"  // FOrdNotEqual (C != is unordered)
                                   : a + " != " + b;
      case VKJIT_BOP_AND: return bl ? a + " && " + b : a + " & " + b;
      case VKJIT_BOP_OR: return bl ? a + " || " + b : a + " | " + b;
      case VKJIT_BOP_XOR: return bl ? a + " != " + b : a + " ^ " + b;
      case VKJIT_BOP_SHL: return s ? "(i32)((u32)" + a + " << ((u32)" + b + " & 31u))" : a + " << (" + b + " & 31u)";
      case VKJIT_BOP_SHR: return s ? a + " >> ((u32)" + b + " & 31u)" : a + " >> (" + b + " & 31u)";
      case VKJIT_BOP_MIN: return f ? "fminf(" + a + ", " + b + ")" : "min(" + a + ", " + b + ")";
      case VKJIT_BOP_MAX: return f ? "fmaxf(" + a + ", " + b + ")" : "max(" + a + ", " + b + ")";
      default: fail(VKJIT_ERR_INVALID, "unknown bop");
    }
  }

  // which vk_math.h entry point: the checked one, or the unchecked fast path when the argument's range proves that the
  // check can never fire
  bool exp_fast(const FRange& a) const { return !no_ranges() && a.ok && a.lo > -87.0f && a.hi < 87.0f; }
  bool log_fast(const FRange& a) const { return !no_ranges() && a.ok && a.lo >= 1.17549435e-38f; }   // positive normal
  bool trig_fast(const FRange& a) const { return !no_ranges() && a.ok && !a.negzero && a.lo >= -105615.0f && a.hi <= 105615.0f; }

  std::string uop_expr(const Var& v, const Val& av) {
    const std::string& a = av.name;
    const TypeId t = v.ty;
    switch (v.kind) {
      case VKJIT_UOP_NEG:
        if (t == VKJIT_TY_F32) return "__uint_as_float(__float_as_uint(" + a + ") ^ 0x80000000u)";
        return t == VKJIT_TY_I32 ? "(i32)(0u - (u32)" + a + ")" : "0u - " + a;
      case VKJIT_UOP_ABS:
        if (t == VKJIT_TY_F32) return "__uint_as_float(__float_as_uint(" + a + ") & 0x7fffffffu)";
        return t == VKJIT_TY_I32 ? "(" + a + " < 0) ? (i32)(0u - (u32)" + a + ") : " + a : a;
      case VKJIT_UOP_NOT: return t == VKJIT_TY_BOOL ? "!" + a : "~" + a;
      case VKJIT_UOP_SQRT: return "__fsqrt_rn(" + a + ")";
      // vk_math.h: the one implementation the oracle compiles too (GPU == oracle bit for bit; <= 1 ulp of the exact value)
      case VKJIT_UOP_EXP: if (fast_math()) return "expf(" + a + ")"; uses_vk_math = true; return (exp_fast(av.fr) ? "vk_expf_fast(" : "vk_expf(") + a + ")";
      case VKJIT_UOP_LOG: if (fast_math()) return "logf(" + a + ")"; uses_vk_math = true; return (log_fast(av.fr) ? "vk_logf_fast(" : "vk_logf(") + a + ")";
      case VKJIT_UOP_SIN: if (fast_math()) return "sinf(" + a + ")"; uses_vk_math = true; return (trig_fast(av.fr) ? "vk_sinf_fast(" : "vk_sinf(") + a + ")";
      case VKJIT_UOP_COS: if (fast_math()) return "cosf(" + a + ")"; uses_vk_math = true; return (trig_fast(av.fr) ? "vk_cosf_fast(" : "vk_cosf(") + a + ")";
      default: fail(VKJIT_ERR_INVALID, "unknown uop");
    }
  }

  std::string cast_expr(TypeId s, TypeId t, const std::string& a) {  // internal.rs:957-992
    if (s == t) return a;
    if (s == VKJIT_TY_U32 && t == VKJIT_TY_I32) return "(i32)" + a;            // bit reinterpret
    if (s == VKJIT_TY_I32 && t == VKJIT_TY_U32) return "(u32)" + a;
    if (s == VKJIT_TY_U32 && t == VKJIT_TY_F32) return "__uint2float_rn(" + a + ")";  // ConvertUToF
    if (s == VKJIT_TY_I32 && t == VKJIT_TY_F32) return "__int2float_rn(" + a + ")";   // ConvertSToF
    if (s == VKJIT_TY_F32 && t == VKJIT_TY_U32) return "__float2uint_rz(" + a + ")";  // intended ConvertFToU
    if (s == VKJIT_TY_F32 && t == VKJIT_TY_I32) return "__float2int_rz(" + a + ")";   // intended ConvertFToS
    if (s == VKJIT_TY_BOOL && t == VKJIT_TY_U32) return a + " ? 1u : 0u";
    if (s == VKJIT_TY_BOOL && t == VKJIT_TY_I32) return a + " ? 1 : 0";
    if (s == VKJIT_TY_BOOL && t == VKJIT_TY_F32) return a + " ? 1.0f : 0.0f";
    if (t == VKJIT_TY_BOOL && s == VKJIT_TY_F32) return a + " != 0.0f";
    if (t == VKJIT_TY_BOOL) return a + " != 0";
    fail(VKJIT_ERR_UNSUPPORTED, "cast");
  }

  // ---- value ranges (see FRange) --------------------------------------------------------------------------------
  static float bits_f(uint32_t w) { float x; memcpy(&x, &w, 4); return x; }
  void node_ranges(uint32_t li) {
    Val& out = vals[li];
    if (!out.elems.empty() || !ty_is_scalar(out.ty)) return;   // struct values carry the ranges of their members
    const Var& v = ir.vars[p.order[li]];
    FRange f; URange u;
    auto D = [&](uint32_t k) -> const Val& { return dep(v, k); };
    switch (v.op) {
      case OP_CONST:
        if (v.ty == VKJIT_TY_F32) f = fr_make(bits_f(v.aux), bits_f(v.aux), v.aux == 0x80000000u);
        else if (v.ty == VKJIT_TY_U32) u.lo = u.hi = v.aux;
        else if (v.ty == VKJIT_TY_BOOL) u.lo = u.hi = v.aux ? 1u : 0u;
        break;
      case OP_ARANGE:   // the lane index: anything below 2^32 (n is a kernel parameter, not part of the key)
        if (v.ty == VKJIT_TY_F32) f = fr_make(0.0f, 4294967296.0f, false);
        break;
      case OP_CAST: {
        const Val& a = D(0);
        if (v.ty == VKJIT_TY_F32 && (a.ty == VKJIT_TY_U32 || a.ty == VKJIT_TY_BOOL))
          f = fr_make((float)a.ur.lo, (float)a.ur.hi, false);       // round-to-nearest is monotone; 0 converts to +0
        else if (v.ty == VKJIT_TY_U32 && a.ty == VKJIT_TY_BOOL) { u.lo = 0u; u.hi = 1u; }
        else if (v.ty == VKJIT_TY_BOOL) { u.lo = 0u; u.hi = 1u; }
        break;
      }
      case OP_BOP: {
        const Val &a = D(0), &b = D(1);
        if (a.ty == VKJIT_TY_F32 && v.ty == VKJIT_TY_F32 && a.fr.ok && b.fr.ok) {
          const FRange &x = a.fr, &y = b.fr;
          switch (v.kind) {
            case VKJIT_BOP_ADD: f = fr_make(x.lo + y.lo, x.hi + y.hi, x.negzero && y.negzero); break;   // x + y = -0 only for (-0) + (-0)
            case VKJIT_BOP_SUB: f = fr_make(x.lo - y.hi, x.hi - y.lo, x.negzero); break;                // only (-0) - (+0)
            case VKJIT_BOP_MUL: {
              const float c[4] = {x.lo * y.lo, x.lo * y.hi, x.hi * y.lo, x.hi * y.hi};
              float lo = c[0], hi = c[0];
              bool fin = true;
              for (float t : c) { fin = fin && f_finite(t); lo = t < lo ? t : lo; hi = t > hi ? t : hi; }
              // a product is +0 or positive when both factors are (and neither is -0); anything else may give -0
              const bool nonneg = x.lo >= 0.0f && !x.negzero && y.lo >= 0.0f && !y.negzero;
              if (fin) f = fr_make(lo, hi, !nonneg);
              break;
            }
            case VKJIT_BOP_MIN: f = fr_make(x.lo < y.lo ? x.lo : y.lo, x.hi < y.hi ? x.hi : y.hi, true); break;
            case VKJIT_BOP_MAX: f = fr_make(x.lo > y.lo ? x.lo : y.lo, x.hi > y.hi ? x.hi : y.hi, true); break;
            default: break;
          }
        } else if (v.ty == VKJIT_TY_U32 && a.ty == VKJIT_TY_U32) {
          const URange &x = a.ur, &y = b.ur;
          auto below_pow2 = [](uint32_t m) { uint32_t r = m; r |= r >> 1; r |= r >> 2; r |= r >> 4; r |= r >> 8; r |= r >> 16; return r; };
          switch (v.kind) {
            case VKJIT_BOP_ADD: if ((uint64_t)x.hi + y.hi <= 0xFFFFFFFFull) { u.lo = x.lo + y.lo; u.hi = x.hi + y.hi; } break;
            case VKJIT_BOP_MUL: if ((uint64_t)x.hi * y.hi <= 0xFFFFFFFFull) { u.lo = x.lo * y.lo; u.hi = x.hi * y.hi; } break;
            case VKJIT_BOP_SHR:   // the generated code shifts by (b & 31)
              if (y.lo == y.hi) { u.lo = x.lo >> (y.lo & 31u); u.hi = x.hi >> (y.lo & 31u); }
              else { u.lo = 0u; u.hi = x.hi; }
              break;
            case VKJIT_BOP_AND: u.lo = 0u; u.hi = x.hi < y.hi ? x.hi : y.hi; break;
            case VKJIT_BOP_OR: case VKJIT_BOP_XOR: u.lo = 0u; u.hi = below_pow2(x.hi | y.hi); break;
            case VKJIT_BOP_MIN: u.lo = x.lo < y.lo ? x.lo : y.lo; u.hi = x.hi < y.hi ? x.hi : y.hi; break;
            case VKJIT_BOP_MAX: u.lo = x.lo > y.lo ? x.lo : y.lo; u.hi = x.hi > y.hi ? x.hi : y.hi; break;
            default: break;
          }
        } else if (v.ty == VKJIT_TY_BOOL) { u.lo = 0u; u.hi = 1u; }
        break;
      }
      case OP_UOP: {
        const Val& a = D(0);
        if (v.ty != VKJIT_TY_F32 || !a.fr.ok || fast_math()) break;
        const FRange& x = a.fr;
        switch (v.kind) {
          case VKJIT_UOP_NEG: f = fr_make(-x.hi, -x.lo, x.lo <= 0.0f && x.hi >= 0.0f); break;
          case VKJIT_UOP_ABS: {
            const float m = std::fabs(x.lo) > std::fabs(x.hi) ? std::fabs(x.lo) : std::fabs(x.hi);
            f = fr_make((x.lo <= 0.0f && x.hi >= 0.0f) ? 0.0f : (std::fabs(x.lo) < std::fabs(x.hi) ? std::fabs(x.lo) : std::fabs(x.hi)), m, false);
            break;
          }
          case VKJIT_UOP_SQRT:   // correctly rounded, monotone; sqrt(-0) = -0
            if (x.lo >= 0.0f) f = fr_make(std::sqrt(x.lo > 0.0f ? x.lo : 0.0f), std::sqrt(x.hi), x.negzero);
            break;
          // vk_math.h results are within 1 ulp of the exact value (exhaustively checked, profiles/r02_vk_math_ulp.md);
          // the bounds below leave 4 ulp plus the rounding of the host's double-precision evaluation
          case VKJIT_UOP_LOG:
            if (x.lo >= 1.17549435e-38f) {
              float lo = f_down((float)std::log((double)x.lo), 4), hi = f_up((float)std::log((double)x.hi), 4);
              if (x.hi <= 1.0f) hi = 0.0f;   // log(x <= 1) is +0 (x = 1) or negative: an error of 1 ulp cannot cross zero
              if (x.lo >= 1.0f) lo = 0.0f;
              f = fr_make(lo, hi, false);    // vk_logf(1) = +0
            }
            break;
          case VKJIT_UOP_EXP:
            if (x.hi < 88.0f) f = fr_make(0.0f, f_up((float)std::exp((double)x.hi), 4), false);
            break;
          case VKJIT_UOP_SIN: case VKJIT_UOP_COS: f = fr_make(-1.00000048f, 1.00000048f, true); break;
          default: break;
        }
        break;
      }
      case OP_SELECT: {
        const Val &a = D(1), &b = D(2);
        if (v.ty == VKJIT_TY_F32 && a.fr.ok && b.fr.ok)
          f = fr_make(a.fr.lo < b.fr.lo ? a.fr.lo : b.fr.lo, a.fr.hi > b.fr.hi ? a.fr.hi : b.fr.hi, a.fr.negzero || b.fr.negzero);
        else if (v.ty == VKJIT_TY_U32 || v.ty == VKJIT_TY_BOOL) {
          u.lo = a.ur.lo < b.ur.lo ? a.ur.lo : b.ur.lo; u.hi = a.ur.hi > b.ur.hi ? a.ur.hi : b.ur.hi;
        }
        break;
      }
      default: break;   // loads, gathers, bit casts, I32 arithmetic: nothing is known
    }
    if (v.op == OP_GETATTR || v.op == OP_SETATTR || v.op == OP_STRUCTINIT || v.op == OP_SCATTER || v.op == OP_SCATTER_ADD) return;  // copies of other values: keep theirs
    out.fr = f; out.ur = u;
  }

  void emit_node(uint32_t li) {
    emit_node_text(li);
    node_ranges(li);
    // bin-range passes: everything after the index of the scatter_add only runs for the lanes of this pass
    if (p.privatize == 3 && (int)li == p.sadd_idx_node)
      line("if ((u32)" + vals[li].name + " - bin_lo >= kbins) return false;");
  }

  void emit_node_text(uint32_t li) {
    const Var& v = ir.vars[p.order[li]];
    const std::string me = "v" + std::to_string(li);
    switch (v.op) {
      case OP_CONST: {
        std::string lit;
        switch (v.ty) {
          case VKJIT_TY_BOOL: lit = v.aux ? "true" : "false"; break;
          case VKJIT_TY_U32: lit = std::to_string(v.aux) + "u"; break;
          case VKJIT_TY_I32: lit = "(i32)" + hex32(v.aux); break;
          default: lit = "__uint_as_float(" + hex32(v.aux) + ")"; break;
        }
        def(li, v.ty, lit);
        break;
      }
      case OP_ARANGE: {  // internal.rs:1078-1094
        const char* idx = v.sharded ? "gi" : "li";
        if (v.ty == VKJIT_TY_U32) def(li, v.ty, idx);
        else if (v.ty == VKJIT_TY_I32) def(li, v.ty, std::string("(i32)") + idx);
        else def(li, v.ty, std::string("__uint2float_rn(") + idx + ")");
        break;
      }
      case OP_BINDING: {
        const Param& pr = p.params[v.aux];
        if (pr.use & USE_STREAM) def(li, v.ty, from_word(v.ty, "in" + std::to_string(v.aux)));
        else { vals[li].name = "/*ptr*/"; vals[li].ty = v.ty; }
        break;
      }
      case OP_BOP: def(li, v.ty, bop_expr(v, dep(v, 0).name, dep(v, 1).name)); break;
      case OP_UOP: {
        const int partner = trig_partner[li];
        if (partner >= 0) {
          if (partner > (int)li) {  // first of the pair: emit both values
            const std::string other = "v" + std::to_string(partner);
            const bool i_am_sin = v.kind == VKJIT_UOP_SIN;
            line("f32 " + me + ", " + other + ";");
            // one range reduction for both; bit-identical to vk_sinf / vk_cosf called separately (checked over all 2^32 inputs)
            line(std::string(fast_math() ? "sincosf(" : trig_fast(dep(v, 0).fr) ? "vk_sincosf_fast(" : "vk_sincosf(") + dep(v, 0).name + ", &" + (i_am_sin ? me : other) + ", &" + (i_am_sin ? other : me) + ");");
            if (!fast_math()) uses_vk_math = true;
          }
          vals[li].name = me; vals[li].ty = v.ty;
          break;
        }
        def(li, v.ty, uop_expr(v, dep(v, 0)));
        break;
      }
      case OP_CAST: def(li, v.ty, cast_expr(dep(v, 0).ty, v.ty, dep(v, 0).name)); break;
      case OP_BITCAST: def(li, v.ty, from_word(v.ty, to_word(dep(v, 0).ty, dep(v, 0).name))); break;
      case OP_GETATTR: vals[li] = dep(v, 0).elems.at(v.aux); break;
      case OP_SETATTR: {  // deps = [src, dst]
        Val r = dep(v, 1);
        r.elems.at(v.aux) = dep(v, 0);
        vals[li] = r;
        break;
      }
      case OP_STRUCTINIT: {
        Val r; r.ty = v.ty;
        for (uint32_t k = 0; k < v.ndeps; ++k) r.elems.push_back(dep(v, k));
        vals[li] = r;
        break;
      }
      case OP_SELECT: {
        int counter = 0;
        vals[li] = select_val(dep(v, 0).name, dep(v, 1), dep(v, 2), me, counter);
        break;
      }
      case OP_GATHER: {  // out = active ? src[idx] : 0 (the reference lowering is broken)
        const uint32_t ps = ir.vars[v.deps()[0]].aux;
        const std::string ld = "g" + std::to_string(ps) + "[(u32)" + dep(v, 1).name + "]";
        const std::string w = v.ndeps >= 3 ? "(" + dep(v, 2).name + " ? " + ld + " : 0u)" : ld;
        def(li, v.ty, from_word(v.ty, w));
        break;
      }
      case OP_SCATTER: case OP_SCATTER_ADD: {  // internal.rs:1056-1077
        const uint32_t pd = ir.vars[v.side_effect].aux;
        const Val& src = dep(v, 0);
        const std::string at = "g" + std::to_string(pd) + " + (u32)" + dep(v, 1).name;
        std::string stmt;
        if (v.op == OP_SCATTER) stmt = "*(" + at + ") = " + to_word(v.ty, src.name) + ";";
        else if (v.ty != VKJIT_TY_F32) {
          // integer scatter_add (mod 2^32 for U32 and I32 alike): warp-aggregated when the warp's lanes collide
          // (kSaddHelper), into the privatised shared-memory bins [0, kbins) where the variant keeps them
          uses_agg = true;
          const std::string act = v.ndeps >= 3 ? dep(v, 2).name : std::string("true");
          const bool priv = p.privatize && (int)pd == p.sadd_param;
          if (priv && p.privatize == 3)   // every lane that gets here has its bin inside this pass's shared-memory range
            line("vk_sadd(g" + std::to_string(pd) + " + bin_lo, vk_sbins, kbins, (u32)" + dep(v, 1).name + " - bin_lo, " + to_word(v.ty, src.name) + ", " + act + ", vk_agg);");
          else
          line("vk_sadd(g" + std::to_string(pd) + ", " + (priv ? "vk_sbins, kbins" : "(u32*)0, 0u") + ", (u32)" + dep(v, 1).name + ", " +
               to_word(v.ty, src.name) + ", " + act + ", vk_agg);");
          vals[li] = src;  // value of the scatter var = src (internal.rs:1076)
          break;
        }
        else if (p.privatize && (int)pd == p.sadd_param) {
          // bins [0, kbins) live in this CTA's shared memory (flushed once at the end of the kernel),
          // the rest go to L2 as before: shared-memory atomics and L2 REDs run side by side
          const std::string ix = "(u32)" + dep(v, 1).name + (p.privatize == 3 ? " - bin_lo" : "");
          if (p.privatize == 3)
            stmt = "atomicAdd(reinterpret_cast<f32*>(vk_sbins + (" + ix + ")), " + src.name + ");";
          else
          stmt = "{ const u32 ix_ = " + ix + "; if (ix_ < kbins) atomicAdd(reinterpret_cast<f32*>(vk_sbins + ix_), " + src.name +
                 "); else atomicAdd(reinterpret_cast<f32*>(g" + std::to_string(pd) + " + ix_), " + src.name + "); }";
        }
        else stmt = "atomicAdd(reinterpret_cast<f32*>(" + at + "), " + src.name + ");";
        if (v.ndeps >= 3) stmt = "if (" + dep(v, 2).name + ") { " + stmt + " }";
        line(stmt);
        vals[li] = src;  // value of the scatter var = src (internal.rs:1076)
        break;
      }
      default: fail(VKJIT_ERR_INVALID, "cannot lower a freed var");
    }
  }
};

}  // namespace

namespace {

// Reduction epilogue for fused trace -> reduce kernels (hand-written; prepended to the generated
// source).  Same structure as prims.cu: warp shuffle -> shared-memory tree -> one partial per CTA ->
// the last CTA to take a ticket folds the partials in a fixed order and resets the ticket.
const char* kReduceEpilogue = R"CUDA(
__device__ __forceinline__ acc_t vk_warp_reduce(acc_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = VK_APPLY(v, VK_FROM_WORD(__shfl_xor_sync(0xFFFFFFFFu, VK_TO_WORD(v), o)));
  return v;
}
__device__ __forceinline__ acc_t vk_block_reduce(acc_t v, acc_t* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = vk_warp_reduce(v);
  if (lane == 0) smem[warp] = v;
  __syncthreads();
  if (warp == 0) {
    acc_t w = lane < (int)(blockDim.x >> 5) ? smem[lane] : VK_IDENTITY;
    w = vk_warp_reduce(w);
    if (lane == 0) smem[0] = w;
  }
  __syncthreads();
  acc_t r = smem[0];
  __syncthreads();
  return r;
}
__device__ __forceinline__ void vk_finish(acc_t acc, u32* partials, unsigned int* ticket, u32* out) {
  __shared__ acc_t smem[32];
  __shared__ bool is_last;
  acc = vk_block_reduce(acc, smem);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = VK_TO_WORD(acc);
    u32 t;  // acquire-release: orders the partial store before it and the last CTA's partial loads after it
    asm volatile("atom.acq_rel.gpu.global.add.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
    is_last = t == gridDim.x - 1;
  }
  __syncthreads();
  if (is_last) {
    acc_t f = VK_IDENTITY;
    for (u32 i = threadIdx.x; i < gridDim.x; i += blockDim.x) f = VK_APPLY(f, VK_FROM_WORD(__ldcg(partials + i)));
    f = vk_block_reduce(f, smem);
    if (threadIdx.x == 0) { out[0] = VK_TO_WORD(f); *ticket = 0u; }
  }
}
)CUDA";

// Integer scatter_add with WARP AGGREGATION (north_star: "scatter-add (warp-aggregated atomics)").  Lanes of a warp that
// hit the same bin are found with match.any, their values summed with redux.sync, and the lowest lane of each group
// issues ONE atomic.  That only pays when lanes collide: 2^26 uniform indices over 2^16 bins almost never do, and
// MATCH + REDUX would be pure overhead on a kernel that is bound by LSU issue.  So every warp PROBES on its first
// scatter_add (st = 2): more than a quarter of its active lanes sharing a bin with another lane -> aggregate from then
// on (st = 1), else plain atomics (st = 0).  Hot-bin inputs (few distinct bins) go from one serialised L2 / shared-
// memory atomic per lane to one per distinct bin per warp.  `sbins`: the privatised shared-memory bins [0, kbins) of
// the variant that keeps part of the target per CTA, or null.
const char* kSaddHelper = R"CUDA(
#if VK_SADD_CLUSTER
// bins [0, kbins) are split over the two CTAs of the cluster: CTA r holds [r * half, (r + 1) * half) in its shared memory
__device__ __forceinline__ u32 vk_cluster_rank() { u32 r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void vk_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ u32 vk_half(const u32 kbins) { return (((kbins + 1u) >> 1) + 3u) & ~3u; }
__device__ __forceinline__ void vk_sadd_one(u32* __restrict__ g, u32* __restrict__ sbins, const u32 kbins, const u32 ix, const u32 val) {
  if (ix < kbins) {
    const u32 half = vk_half(kbins), owner = ix >= half ? 1u : 0u;
    const u32 local = (u32)__cvta_generic_to_shared(sbins + (ix - owner * half));
    u32 remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(owner));
    asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(val) : "memory");
  } else atomicAdd(g + ix, val);
}
#else
__device__ __forceinline__ void vk_sadd_one(u32* __restrict__ g, u32* __restrict__ sbins, const u32 kbins, const u32 ix, const u32 val) {
  if (ix < kbins) atomicAdd(sbins + ix, val);
  else atomicAdd(g + ix, val);
}
#endif
__device__ __forceinline__ void vk_sadd(u32* __restrict__ g, u32* __restrict__ sbins, const u32 kbins, const u32 ix, const u32 val,
                                        const bool active, u32& st) {
  if (st == 0u) {
    if (active) vk_sadd_one(g, sbins, kbins, ix, val);
    return;
  }
  const unsigned vote = __ballot_sync(__activemask(), active);
  if (!active) return;
  const unsigned peers = __match_any_sync(vote, ix);   // the active lanes of this warp that target bin ix
  // the decision is taken by ALL lanes of `vote` together whenever one of them still probes (lanes that were masked off
  // at the warp's first scatter_add join later with st == 2): a ballot under `if (st == 2u)` alone would name lanes in
  // its mask that never execute it
  if (__ballot_sync(vote, st == 2u) != 0u) {
    const unsigned dup = __ballot_sync(vote, (peers & (peers - 1u)) != 0u);
    st = (__popc(dup) * 4 > __popc(vote)) ? 1u : 0u;
  }
  const u32 total = __reduce_add_sync(peers, val);
  if ((peers & ((1u << (threadIdx.x & 31)) - 1u)) == 0u) vk_sadd_one(g, sbins, kbins, ix, total);
}
)CUDA";

// scan_common.cuh / scan_fused.cuh as text (generated by the Makefile)
#include "build/scan_src.inc"

// Shell of a fused trace -> scan kernel: the pointer block, the adapter from the ring's words to vk_lane, and
// the hand-written kernel text specialised through three macros.
std::string scan_shell(const Program& p, const std::vector<uint32_t>& streams, const std::vector<uint32_t>& ptrs, size_t nroots) {
  const size_t ns = streams.size();
  if (ns > (size_t)kScanFusedMaxStreams) fail(VKJIT_ERR_UNSUPPORTED, "fused scan: too many streamed arrays");
  if (nroots != (p.scan == SCAN_COMPRESS_VALUE ? 2u : 1u)) fail(VKJIT_ERR_INVALID, "fused scan: wrong number of roots");
  std::string s;
  const ScanFusedGeom geom = scan_fused_geom(ns, p.scan, p.order.size());
  s += "#define VK_SCAN_MODE " + std::to_string(p.scan) + "\n#define VK_NS " + std::to_string(ns) + "\n#define VK_VPT " +
       std::to_string(geom.vpt) + "\n#define VK_LAG " + std::to_string(geom.lag ? 1 : 0) + "\n#define VK_SLOTS " +
       std::to_string(geom.slots) + "\n#define VK_T " + std::to_string(geom.threads) + "\n#define VK_LOOK_WIDE " + std::to_string(geom.look_wide) + "\n#define VK_TRACE " + (fscan_trace_file() ? "1" : "0") + "\n#define VK_EARLY " + (scan_early() ? "1" : "0") + "\n#define VK_WREG " + (scan_wreg() ? "1" : "0") + "\n#define VK_PARK " + (geom.park ? "1" : "0") + "\n#define VK_CTRL " + (geom.ctrl ? "1" : "0") +
       "\n#define VK_DEPTH " + std::to_string(geom.depth) + "\n#define VK_CTAS " + std::to_string(geom.ctas) + "\n#define VK_CLAG " + std::to_string(geom.clag) + "\n#define VK_DIAG " + std::to_string(scan_diag() & 1) + "\n#define VK_COALESCE " + ((scan_diag() & 2) ? "1" : "0") + "\n#define VK_PACKED " + ((scan_diag() & 4) ? "0" : "1") + "\n#define VK_LAGPACK " + ((geom.lag && !geom.ctrl && scan_lag_packed()) ? "1" : "0") + "\n";
  s += "struct VkPtrs {\n  const u32* s[" + std::to_string(std::max<size_t>(ns, 1)) + "];  // streamed arrays (staged by TMA)\n";
  for (uint32_t k : ptrs) {
    if (p.params[k].use & USE_SCATTER) fail(VKJIT_ERR_UNSUPPORTED, "fused scan: the trace has side effects");
    s += "  const u32* g" + std::to_string(k) + ";\n";
  }
  s += "};\n";
  s += "__device__ __forceinline__ void vk_eval(const VkPtrs& P, const u32 gi, const u32 li, const u32* in, u32& o0, u32& o1) {\n  vk_lane(gi, li";
  for (size_t i = 0; i < ns; ++i) s += ", in[" + std::to_string(i) + "]";
  s += ", o0";
  if (nroots == 2) s += ", o1";
  for (uint32_t k : ptrs) s += ", P.g" + std::to_string(k);
  s += ");\n}\n\n";
  s += kScanCommonSrc;
  s += kScanFusedSrc;
  return s;
}

std::string reduce_defines(int red, TypeId ty) {
  std::string d = std::string("typedef ") + ctype(ty) + " acc_t;\n";
  d += "#define VK_FROM_WORD(w) (" + from_word(ty, "(w)") + ")\n";
  d += "#define VK_TO_WORD(v) (" + to_word(ty, "(v)") + ")\n";
  const bool f = ty == VKJIT_TY_F32, s = ty == VKJIT_TY_I32;
  std::string apply, ident;
  switch (red) {
    case VKJIT_RED_SUM:
      apply = f ? "__fadd_rn((a), (b))" : s ? "(i32)((u32)(a) + (u32)(b))" : "((a) + (b))";
      ident = f ? "0.0f" : s ? "0" : "0u";
      break;
    case VKJIT_RED_MIN:
      apply = f ? "fminf((a), (b))" : "min((a), (b))";
      ident = f ? "__uint_as_float(0x7fc00000u)" : s ? "(i32)0x7fffffffu" : "0xffffffffu";
      break;
    default:
      apply = f ? "fmaxf((a), (b))" : "max((a), (b))";
      ident = f ? "__uint_as_float(0x7fc00000u)" : s ? "(i32)0x80000000u" : "0u";
      break;
  }
  d += "#define VK_APPLY(a, b) (" + apply + ")\n#define VK_IDENTITY (" + ident + ")\n";
  return d;
}

}  // namespace

// Fingerprint of everything that turns a canonical key into a cubin besides the key itself: the embedded device
// sources (vk_math.h, scan_common.cuh, scan_fused.cuh, the reduction epilogue) and a generator revision that is bumped
// whenever generate_cuda's output or a kernel's parameter layout changes.  The on-disk cubin cache stores it: a cubin
// written by another build of the library is a miss, never a kernel with the wrong argument ABI.
uint32_t generator_fingerprint() {
  static const uint32_t fp = [] {
    constexpr uint32_t kGeneratorRevision = 13;  // round 2: vk_math.h lowering, fast-math variant bit, range-proven fast paths
    uint32_t h = 2166136261u ^ kGeneratorRevision;
    auto mix = [&](const char* t) { for (; *t; ++t) { h ^= (unsigned char)*t; h *= 16777619u; } };
    mix(kVkMathSrc); mix(kScanCommonSrc); mix(kScanFusedSrc); mix(kReduceEpilogue); mix(kSaddHelper);
    return h;
  }();
  return fp;
}

std::string generate_cuda(const Ir& ir, const Program& p) {
  const bool reduce = p.reduce >= 0;
  Gen g(ir, p);
  for (uint32_t li = 0; li < p.order.size(); ++li) g.emit_node(li);
  for (size_t r = 0; r < p.roots.size(); ++r) {
    const Val& v = g.vals[p.roots[r]];
    // (the proven range travels as a comment: tests/test_product_cpu.py checks it against the oracle's values)
    std::string note;
    if (v.elems.empty() && v.ty == VKJIT_TY_F32 && v.fr.ok) {
      uint32_t lo, hi; memcpy(&lo, &v.fr.lo, 4); memcpy(&hi, &v.fr.hi, 4);
      note = "  // range f32 " + hex32(lo) + " " + hex32(hi) + (v.fr.negzero ? " negzero" : " no-negzero");
    } else if (v.elems.empty() && v.ty == VKJIT_TY_U32 && !(v.ur.lo == 0u && v.ur.hi == 0xFFFFFFFFu)) {
      note = "  // range u32 " + std::to_string(v.ur.lo) + " " + std::to_string(v.ur.hi);
    }
    g.line("out" + std::to_string(r) + " = " + to_word(v.ty, v.name) + ";" + note);
  }

  std::vector<uint32_t> streams, ptrs;
  for (uint32_t k = 0; k < p.params.size(); ++k) {
    if (p.params[k].use & USE_STREAM) streams.push_back(k);
    if (p.params[k].use & (USE_GATHER | USE_SCATTER)) ptrs.push_back(k);
  }
  const size_t nroots = p.roots.size();

  std::string s;
  s += "// vkjit-b200 fused trace kernel; key " + std::to_string(p.hash.lo) + ":" + std::to_string(p.hash.hi) + "\n";
  s += "typedef unsigned int u32;\ntypedef int i32;\ntypedef float f32;\n\n";
  if (g.uses_vk_math) s += std::string(kVkMathSrc) + "\n";
  const bool scan = p.scan >= 0;
  if (scan) s += "typedef unsigned int uint32_t;\ntypedef unsigned long long uint64_t;\n\n";
  if (reduce) {
    if (nroots != 1) fail(VKJIT_ERR_INVALID, "a fused reduction has exactly one root");
    s += reduce_defines(p.reduce, g.vals[p.roots[0]].ty) + kReduceEpilogue + "\n";
  }

  const bool priv = p.privatize != 0;
  bool priv_f32 = false;
  if (priv) {
    priv_f32 = ir.vars[p.params[p.sadd_param].var].ty == VKJIT_TY_F32;
    s += "extern __shared__ u32 vk_sbins[];  // privatised scatter_add bins [0, kbins)\n\n";
  }

  const bool cluster = p.privatize == 2 && !priv_f32 && g.uses_agg;
  if (g.uses_agg) s += std::string("#define VK_SADD_CLUSTER ") + (cluster ? "1" : "0") + "\n" + kSaddHelper + "\n";
  // per-lane body
  const bool passes = p.privatize == 3;   // vk_lane returns whether the lane belongs to this bin-range pass
  s += std::string("__device__ __forceinline__ ") + (passes ? "bool" : "void") + " vk_lane(const u32 gi, const u32 li";
  if (priv) s += ", const u32 kbins";
  if (passes) s += ", const u32 bin_lo";
  if (g.uses_agg) s += ", u32& vk_agg";
  for (uint32_t k : streams) s += ", const u32 in" + std::to_string(k);
  for (size_t r = 0; r < nroots; ++r) s += ", u32& out" + std::to_string(r);
  for (uint32_t k : ptrs) {
    const bool w = p.params[k].use & USE_SCATTER;
    s += std::string(", ") + (w ? "u32* " : "const u32* __restrict__ ") + "g" + std::to_string(k);
  }
  s += ") {\n" + g.body + (passes ? "  return true;\n" : "") + "}\n\n";
  if (scan) return s + scan_shell(p, streams, ptrs, nroots);

  auto call = [&](const std::string& gi, const std::string& li, const char* comp, bool vec) {
    std::string c = "vk_lane(" + gi + ", " + li;
    if (priv) c += ", kbins";
    if (passes) c += ", bin_lo";
    if (g.uses_agg) c += ", vk_agg";
    for (uint32_t k : streams) c += ", a" + std::to_string(k) + (vec ? std::string(".") + comp : "");
    for (size_t r = 0; r < nroots; ++r) c += ", r" + std::to_string(r) + (vec ? std::string(".") + comp : "");
    for (uint32_t k : ptrs) c += ", p" + std::to_string(k);
    return c + ")" + (passes ? "" : ";");
  };

  s += std::string("extern \"C\" __global__ void ") + (cluster ? "__cluster_dims__(2, 1, 1) " : "") + "__launch_bounds__(" + (priv ? "1024" : "256") +
       ") vkjit_trace(const u32 n, const u32 base";
  if (priv) s += ", const u32 kbins";
  if (passes) s += ", const u32 bin_lo";
  for (uint32_t k = 0; k < p.params.size(); ++k) {
    const bool w = p.params[k].use & USE_SCATTER;
    s += std::string(",\n    ") + (w ? "u32* " : "const u32* __restrict__ ") + "p" + std::to_string(k);
  }
  if (reduce) s += ",\n    u32* __restrict__ partials, unsigned int* __restrict__ ticket, u32* __restrict__ o0";
  else for (size_t r = 0; r < nroots; ++r) s += ",\n    u32* __restrict__ o" + std::to_string(r);
  s += ") {\n";
  if (reduce) s += "  acc_t c0 = VK_IDENTITY, c1 = VK_IDENTITY, c2 = VK_IDENTITY, c3 = VK_IDENTITY;\n";
  if (cluster) s += "  for (u32 i = threadIdx.x; i < vk_half(kbins); i += blockDim.x) vk_sbins[i] = 0u;\n  vk_cluster_sync();  // both halves are zero before anyone adds\n";
  else if (priv) s += "  for (u32 i = threadIdx.x; i < kbins; i += blockDim.x) vk_sbins[i] = 0u;\n  __syncthreads();\n";
  if (g.uses_agg) s += std::string("  u32 vk_agg = ") + (no_agg() ? "0u" : "2u") + ";  // warp aggregation of integer scatter_add: 2 probe, 1 aggregate, 0 plain atomics\n";
  s += "  const u32 tid = blockIdx.x * blockDim.x + threadIdx.x;\n";
  s += "  const u32 nthreads = gridDim.x * blockDim.x;\n";
  if (p.vectorized) {
    // 128-bit vectorised, coalesced main loop: 4 consecutive lanes per thread and iteration
    s += "  const u32 nvec = n >> 2;\n";
    s += "#pragma unroll " + std::to_string(unroll_factor()) + "\n";
    s += "  for (u32 v = tid; v < nvec; v += nthreads) {\n";
    for (uint32_t k : streams)
      s += "    const uint4 a" + std::to_string(k) + " = reinterpret_cast<const uint4*>(p" + std::to_string(k) + ")[v];\n";
    for (size_t r = 0; r < nroots; ++r) s += "    uint4 r" + std::to_string(r) + ";\n";
    s += "    const u32 l0 = v << 2;\n";
    const char* comps[4] = {"x", "y", "z", "w"};
    for (int j = 0; j < 4; ++j) {
      if (passes) {  // 128-bit loads, but a lane's roots are written (4 bytes) in the pass that owns its bin
        s += "    if (" + call("base + l0 + " + std::to_string(j) + "u", "l0 + " + std::to_string(j) + "u", comps[j], true) + ") {";
        for (size_t r = 0; r < nroots; ++r) s += " o" + std::to_string(r) + "[l0 + " + std::to_string(j) + "u] = r" + std::to_string(r) + "." + comps[j] + ";";
        s += " }\n";
      } else
      s += "    " + call("base + l0 + " + std::to_string(j) + "u", "l0 + " + std::to_string(j) + "u", comps[j], true) + "\n";
    }
    if (passes) {}
    else if (reduce)
      s += "    c0 = VK_APPLY(c0, VK_FROM_WORD(r0.x)); c1 = VK_APPLY(c1, VK_FROM_WORD(r0.y));\n"
           "    c2 = VK_APPLY(c2, VK_FROM_WORD(r0.z)); c3 = VK_APPLY(c3, VK_FROM_WORD(r0.w));\n";
    else
      for (size_t r = 0; r < nroots; ++r)
        s += "    reinterpret_cast<uint4*>(o" + std::to_string(r) + ")[v] = r" + std::to_string(r) + ";\n";
    s += "  }\n";
    s += "  for (unsigned long long i = (unsigned long long)(nvec << 2) + tid; i < n; i += nthreads) {\n";
  } else {
    s += "  for (unsigned long long i = tid; i < n; i += nthreads) {\n";
  }
  for (uint32_t k : streams) s += "    const u32 a" + std::to_string(k) + " = p" + std::to_string(k) + "[i];\n";
  for (size_t r = 0; r < nroots; ++r) s += "    u32 r" + std::to_string(r) + ";\n";
  if (passes) {  // the lane's roots are written in the pass that owns its bin
    s += "    if (" + call("base + (u32)i", "(u32)i", "", false) + ") {\n";
    for (size_t r = 0; r < nroots; ++r) s += "      o" + std::to_string(r) + "[i] = r" + std::to_string(r) + ";\n";
    s += "    }\n";
  } else {
  s += "    " + call("base + (u32)i", "(u32)i", "", false) + "\n";
  if (reduce) s += "    c0 = VK_APPLY(c0, VK_FROM_WORD(r0));\n";
  else for (size_t r = 0; r < nroots; ++r) s += "    o" + std::to_string(r) + "[i] = r" + std::to_string(r) + ";\n";
  }
  s += "  }\n";
  if (reduce) s += "  vk_finish(VK_APPLY(VK_APPLY(c0, c1), VK_APPLY(c2, c3)), partials, ticket, o0);\n";
  if (priv) {
    const std::string g = "p" + std::to_string(p.sadd_param);
    if (cluster) {
      s += "  vk_cluster_sync();  // every add of both CTAs has landed\n";
      s += "  { const u32 half = vk_half(kbins), lo = vk_cluster_rank() * half;\n";
      s += "    for (u32 i = threadIdx.x; i < half && lo + i < kbins; i += blockDim.x) {\n      const u32 w = vk_sbins[i];\n";
      s += "      if (w != 0u) atomicAdd(" + g + " + lo + i, w);\n    }\n  }\n";
    } else {
      s += "  __syncthreads();\n  for (u32 i = threadIdx.x; i < kbins; i += blockDim.x) {\n    const u32 w = vk_sbins[i];\n";
      const std::string at = passes ? g + " + bin_lo + i" : g + " + i";
      if (priv_f32) s += "    if (w != 0u) atomicAdd(reinterpret_cast<f32*>(" + at + "), __uint_as_float(w));\n";
      else s += "    if (w != 0u) atomicAdd(" + at + ", w);\n";
      s += "  }\n";
    }
  }
  s += "}\n";
  return s;
}

}  // namespace vkjit
