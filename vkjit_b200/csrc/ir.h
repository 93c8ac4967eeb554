// ir.h — trace IR of the B200 backend.
//
// Observable contract follows vkjit_core::Ir (reference libs/vkjit-core/src/internal.rs:105-542):
// same ops, same type promotion, same ref-count bookkeeping as seen by a front-end.  The storage
// is different on purpose: vars live in a flat table with a free list (the reference never
// reuses a slot, internal.rs:206-208), dependencies are stored inline, the array lives on the
// var instead of in a side HashMap, and every var carries scratch fields so that the per-eval
// trace walk needs no hashing or allocation (cached-launch budget: < 10 us).
#pragma once
#include <atomic>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"

namespace vkjit {

enum Op : uint8_t {
  OP_FREE = 0,  // slot on the free list
  // reference op set, internal.rs:64-77
  OP_BINDING, OP_BOP, OP_ARANGE, OP_CONST, OP_GETATTR, OP_SETATTR, OP_STRUCTINIT,
  OP_GATHER, OP_SCATTER, OP_SELECT, OP_CAST,
  // extensions
  OP_UOP, OP_BITCAST, OP_SCATTER_ADD
};

// Device array owned by a Binding var — Backend::Array (backend/mod.rs:8-12).
struct Array {
  std::atomic<uint32_t> refs{1};  // the owning var holds one; every exported DLPack tensor holds one more
  void* ptr = nullptr;
  size_t bytes = 0;     // logical size (Array::size)
  size_t capacity = 0;  // allocation size (compress over-allocates)
  bool owned = true;    // false: view of foreign device memory (vkjit_array_wrap_device), never freed here
  bool exposed = false; // the device pointer was handed out (device_ptr / DLPack / CUDA Array Interface): others may write it
  void (*release)(void*) = nullptr;  // view with an owner (DLPack import): called once, after the Ir lock is dropped
  void* release_ctx = nullptr;
};

// One cache line per var: the per-eval trace walk (program.cpp: build_program, the cache-hit critical path) touches
// every reachable var once, and everything it reads sits in the first 40 bytes.
struct alignas(64) Var {  // internal.rs:105-114
  Op op = OP_FREE;
  uint8_t sharded = 0;   // lanes are this rank's slice of a global 1-D range
  uint16_t kind = 0;     // Bop / Uop kind
  TypeId ty = VKJIT_TY_VOID;
  uint32_t aux = 0;      // Const bit pattern | GetAttr/SetAttr index | (during a walk) a Binding's param index
  uint32_t ndeps = 0;
  union {                // a var with dependencies is never an Arange or a Binding, so the two views share storage
    struct {
      VarId dep_inline[3];
      VarId side_effect;   // Scatter target (side_effects[0], internal.rs:396)
    };
    struct {
      uint64_t num;        // Arange(n)
      uint64_t base;       // first global lane of a sharded Arange / of a ragged sharded Binding
    };
  };
  // scratch for trace walks
  uint32_t stamp = 0;
  uint32_t local = 0;
  uint32_t ref_count = 0;
  bool has_se = false;
  std::unique_ptr<VarId[]> dep_ext;  // StructInit with more than 3 members
  Array* array = nullptr;            // replaces Ir.arrays: HashMap<VarId, Array> (internal.rs:130)

  Var() : num(0), base(0) {}
  const VarId* deps() const { return ndeps <= 3 ? dep_inline : dep_ext.get(); }
};
static_assert(sizeof(Var) == 64, "Var is meant to be exactly one cache line");

class Ir {
 public:
  Ir();
  ~Ir();
  Ir(const Ir&) = delete;

  std::mutex mu;  // the front-ends' `lazy_static IR: Mutex<Ir>` (vkjit-rust/src/lib.rs:9-11)
  std::vector<Var> vars;
  std::vector<VarId> free_list;
  std::vector<VarId> schedule;
  std::vector<std::vector<TypeId>> struct_types;
  size_t n_arrays = 0;
  uint32_t stamp_counter = 0;

  Var& var(VarId id);
  const Var& var(VarId id) const;
  bool is_buffer(VarId id) const { return id < vars.size() && vars[id].array != nullptr; }

  // types
  TypeId struct_type(const TypeId* elems, size_t n);
  const std::vector<TypeId>& struct_elems(TypeId t) const;
  void check_type(TypeId t) const;
  static TypeId ty_max(TypeId a, TypeId b);  // derive(Ord) max, vartype.rs:24-33

  // constructors (internal.rs:218-400)
  VarId constant(TypeId ty, uint32_t bits);
  VarId binding(TypeId ty, Array* arr, bool sharded);
  VarId arange(TypeId ty, uint64_t n, uint64_t base, bool sharded);
  VarId linspace(TypeId ty, VarId start, VarId stop, uint64_t n);
  VarId zeros(TypeId ty);
  VarId ones(TypeId ty);
  VarId cast(VarId src, TypeId ty);
  VarId bop(int kind, VarId lhs, VarId rhs);
  VarId uop(int kind, VarId src);
  VarId bitcast(VarId src, TypeId ty);
  VarId select(VarId c, VarId l, VarId r);
  VarId struct_init(const VarId* elems, size_t n);
  VarId getattr(VarId src, size_t idx);
  VarId setattr(VarId dst, VarId src, size_t idx);
  VarId gather(VarId src, VarId idx, bool has_active, VarId active);
  VarId scatter(Op op, VarId src, VarId dst, VarId idx, bool has_active, VarId active);

  // lifetime (internal.rs:450-481)
  void inc_ref(VarId id);
  void dec_ref(VarId id);
  void do_schedule(const VarId* ids, size_t n);
  void clear_schedule();

  // after a kernel ran: roots become Bindings owning `outs` (internal.rs:492-521); `order` = post-order of the
  // trace that ran (Program::order), which is what makes the release a single reverse sweep
  void commit_roots(const std::vector<VarId>& roots, const std::vector<Array*>& outs, const std::vector<VarId>& order);

  // repr
  std::string repr() const;                // {:#?} of the Ir
  std::string var_debug(VarId id) const;   // {:?} of one Var
  std::string type_name(TypeId t) const;

  uint32_t next_stamp();

 private:
  VarId alloc_slot();
  VarId new_var(Op op, TypeId ty, const VarId* deps, size_t ndeps, uint16_t kind = 0, uint32_t aux = 0);
  std::vector<VarId> dec_stack_;
};

// Rust `{:?}` text of an f32 (shortest round-trip digits, always a fractional part).
std::string format_f32(float f);

// Implemented by the runtime: returns the memory of a dying Binding to the pool.  Owner callbacks of foreign
// views are queued per thread and run by drain_foreign_releases() once the caller holds no Ir lock (an owner
// may be one of our own exported DLPack tensors, whose deleter takes that lock).
void release_array(Array* a);   // drops one reference; the last one frees
inline void retain_array(Array* a) { a->refs.fetch_add(1, std::memory_order_relaxed); }
void drain_foreign_releases();

}  // namespace vkjit
