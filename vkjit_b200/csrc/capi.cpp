// capi.cpp — extern "C" surface declared in include/vkjit_b200.h.
// Every entry point turns the reference's panics into a status + thread-local message.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>

#include "dist.h"
#include "ir.h"
#include "prims.h"
#include "program.h"
#include "runtime.h"

using namespace vkjit;

struct vkjit_ir {
  Ir ir;
};

namespace {

thread_local std::string g_last_error;

template <class F>
vkjit_status guard(F&& f) {
  try {
    f();
    return VKJIT_OK;
  } catch (const Error& e) {
    g_last_error = e.what();
    return e.code;
  } catch (const std::exception& e) {
    g_last_error = e.what();
    return VKJIT_ERR_INVALID;
  }
}

template <class F>
vkjit_status with_ir(vkjit_ir* h, F&& f) {
  if (!h) { g_last_error = "null vkjit_ir"; return VKJIT_ERR_INVALID; }
  const vkjit_status st = guard([&] {
    std::lock_guard<std::mutex> lock(h->ir.mu);
    f(h->ir);
  });
  drain_foreign_releases();  // owners of dropped foreign views, outside the lock
  return st;
}

vkjit_status copy_out(const std::string& s, char* buf, size_t cap, size_t* out_len) {
  if (out_len) *out_len = s.size();
  if (buf && cap) {
    const size_t m = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), m);
    buf[m] = 0;
  }
  return VKJIT_OK;
}

// Ir::array_* (internal.rs:313-348): fresh device array + H2D copy
VarId upload(Ir& ir, TypeId ty, const void* data, size_t n, bool sharded) {
  Backend& be = Backend::get();
  Array* a = be.new_array(n * 4);
  try {
    if (n) be.h2d(a->ptr, data, n * 4);
  } catch (...) { release_array(a); throw; }
  return ir.binding(ty, a, sharded);
}

bool misaligned(const Ir& ir, VarId id) {
  return ir.is_buffer(id) && (((uintptr_t)ir.var(id).array->ptr) & 15u) != 0;
}

// The hand-written primitives use 128-bit / TMA accesses: an unevaluated var is evaluated, and a view of foreign
// memory that is not 16-byte aligned is first copied into pool memory (eval of a Binding root copies it,
// internal.rs:1192-1205; the scalar kernel variant does the copy).
void ensure_buffer(Ir& ir, VarId id) {
  if (!ir.is_buffer(id) || misaligned(ir, id)) eval(ir, &id, 1);
}

// Does the trace below `id` contain a Scatter / ScatterAdd?  The eager primitives evaluate an unevaluated operand
// inside their own kernel (or into a temporary) WITHOUT committing it; a scatter in that trace would run there and
// again at the var's own eval.  Such an operand is committed first — eval([id]), the side effect runs exactly once,
// the var becomes a buffer — and the hand-written primitive reads the buffer (same rule in the oracle).
bool has_side_effects(Ir& ir, VarId id) {
  if (ir.is_buffer(id)) return false;
  static thread_local Program p;
  std::vector<VarId> roots{id};
  build_program(ir, roots, true, p);
  if (p.n == 0) return false;  // (an empty shard has nothing to run)
  for (const Param& pr : p.params)
    if (pr.use & USE_SCATTER) return true;
  return false;
}
// Operands of ONE primitive call are committed by ONE eval: a scatter they share still runs once.
void commit_if_side_effects(Ir& ir, const VarId* ids, size_t n) {
  std::vector<VarId> roots;
  for (size_t i = 0; i < n; ++i)
    if (has_side_effects(ir, ids[i])) roots.push_back(ids[i]);
  if (!roots.empty()) eval(ir, roots.data(), roots.size());
}

// Buffer::str (internal.rs:404-422) through a D2H copy
std::string buffer_str(Ir& ir, VarId id) {
  const Var& v = ir.var(id);
  const size_t n = v.array->bytes / 4;
  std::vector<uint32_t> w(n);
  Backend::get().d2h(w.data(), v.array->ptr, n * 4);
  std::string o = "[";
  auto sep = [&](size_t i) { if (i) o += ", "; };
  for (size_t i = 0; i < n; ++i) {
    if (v.ty == VKJIT_TY_BOOL) {  // printed as raw u8 (internal.rs:417-419): 4 bytes per element
      const uint8_t* b = (const uint8_t*)&w[i];
      for (int k = 0; k < 4; ++k) { sep(i * 4 + k); o += std::to_string(b[k]); }
      continue;
    }
    sep(i);
    if (v.ty == VKJIT_TY_U32) o += std::to_string(w[i]);
    else if (v.ty == VKJIT_TY_I32) o += std::to_string((int32_t)w[i]);
    else if (v.ty == VKJIT_TY_F32) o += format_f32(bits_f32(w[i]));
    else return "Undefined Type!";
  }
  return o + "]";
}

}  // namespace

extern "C" {

// ---- lifecycle ------------------------------------------------------------------------------
vkjit_status vkjit_init(int32_t device) { return guard([&] { Backend::init(device); }); }
vkjit_status vkjit_shutdown(void) { return guard([&] { dist::shutdown(); Backend::shutdown(); }); }
int32_t vkjit_is_initialized(void) { return Backend::initialized() ? 1 : 0; }
const char* vkjit_last_error(void) { return g_last_error.c_str(); }
uint32_t vkjit_abi_version(void) { return VKJIT_B200_ABI_VERSION; }
vkjit_status vkjit_stream(void** out) { return guard([&] { *out = Backend::get().wait_stream(); }); }
vkjit_status vkjit_device(int32_t* out) { return guard([&] { *out = Backend::get().device; }); }
vkjit_status vkjit_sync(void) { return guard([&] { Backend::get().sync(); }); }
vkjit_status vkjit_host_alloc(size_t bytes, void** out) {
  return guard([&] {
    Backend::get();
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 16);
    if (e != cudaSuccess) fail(VKJIT_ERR_CUDA, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
  });
}
vkjit_status vkjit_host_free(void* p) {
  return guard([&] { if (p) cudaFreeHost(p); });
}

vkjit_status vkjit_ir_create(vkjit_ir** out) { return guard([&] { *out = new vkjit_ir(); }); }
vkjit_status vkjit_ir_destroy(vkjit_ir* h) {
  const vkjit_status st = guard([&] { delete h; });
  drain_foreign_releases();
  return st;
}

// ---- types ------------------------------------------------------------------------------------
vkjit_status vkjit_type_struct(vkjit_ir* h, const vkjit_type* e, size_t n, vkjit_type* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.struct_type(e, n); });
}
vkjit_status vkjit_type_struct_len(vkjit_ir* h, vkjit_type t, size_t* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.struct_elems(t).size(); });
}
vkjit_status vkjit_type_struct_elem(vkjit_ir* h, vkjit_type t, size_t i, vkjit_type* out) {
  return with_ir(h, [&](Ir& ir) {
    const auto& e = ir.struct_elems(t);
    if (i >= e.size()) fail(VKJIT_ERR_INVALID, "struct member index out of range");
    *out = e[i];
  });
}

// ---- constructors -------------------------------------------------------------------------------
vkjit_status vkjit_const_f32(vkjit_ir* h, float v, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.constant(VKJIT_TY_F32, f32_bits(v)); }); }
vkjit_status vkjit_const_i32(vkjit_ir* h, int32_t v, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.constant(VKJIT_TY_I32, (uint32_t)v); }); }
vkjit_status vkjit_const_u32(vkjit_ir* h, uint32_t v, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.constant(VKJIT_TY_U32, v); }); }
vkjit_status vkjit_const_bool(vkjit_ir* h, int32_t v, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.constant(VKJIT_TY_BOOL, v ? 1u : 0u); }); }
vkjit_status vkjit_array_f32(vkjit_ir* h, const float* d, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = upload(ir, VKJIT_TY_F32, d, n, false); }); }
vkjit_status vkjit_array_i32(vkjit_ir* h, const int32_t* d, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = upload(ir, VKJIT_TY_I32, d, n, false); }); }
vkjit_status vkjit_array_u32(vkjit_ir* h, const uint32_t* d, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = upload(ir, VKJIT_TY_U32, d, n, false); }); }
vkjit_status vkjit_array_bool(vkjit_ir* h, const uint32_t* d, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = upload(ir, VKJIT_TY_BOOL, d, n, false); }); }
vkjit_status vkjit_array_empty(vkjit_ir* h, vkjit_type ty, size_t n, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_scalar(ty)) fail(VKJIT_ERR_TYPE, "array of a non-scalar type");
    *out = ir.binding(ty, Backend::get().new_array(n * 4), false);
  });
}
vkjit_status vkjit_array_wrap_device(vkjit_ir* h, vkjit_type ty, uint64_t device_ptr, size_t n, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_scalar(ty)) fail(VKJIT_ERR_TYPE, "array of a non-scalar type");
    if (!device_ptr && n) fail(VKJIT_ERR_INVALID, "null device pointer");
    if (device_ptr & 3u) fail(VKJIT_ERR_INVALID, "device pointer must be 4-byte aligned");
    // no backend needed: the view is only dereferenced by kernels (trace construction and code generation
    // over views work without a device; eval fails loudly there)
    Array* a = new Array();
    a->ptr = (void*)(uintptr_t)device_ptr; a->bytes = n * 4; a->capacity = n * 4; a->owned = false;
    *out = ir.binding(ty, a, false);
  });
}
vkjit_status vkjit_array_wrap_device_owned(vkjit_ir* h, vkjit_type ty, uint64_t device_ptr, size_t n, void (*release)(void*),
                                           void* ctx, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_scalar(ty)) fail(VKJIT_ERR_TYPE, "array of a non-scalar type");
    if (!device_ptr && n) fail(VKJIT_ERR_INVALID, "null device pointer");
    if (device_ptr & 3u) fail(VKJIT_ERR_INVALID, "device pointer must be 4-byte aligned");
    Array* a = new Array();
    a->ptr = (void*)(uintptr_t)device_ptr; a->bytes = n * 4; a->capacity = n * 4; a->owned = false;
    try { *out = ir.binding(ty, a, false); } catch (...) { delete a; throw; }
    a->release = release; a->release_ctx = ctx;  // only now: on failure the caller still owns ctx
  });
}

// ---- DLPack (dlpack.h v0.8 "dltensor" ABI) ---------------------------------------------------------
namespace {
struct DLDevice { int32_t device_type, device_id; };
struct DLDataType { uint8_t code, bits; uint16_t lanes; };
struct DLTensor { void* data; DLDevice device; int32_t ndim; DLDataType dtype; int64_t* shape; int64_t* strides; uint64_t byte_offset; };
struct DLManagedTensor { DLTensor dl_tensor; void* manager_ctx; void (*deleter)(DLManagedTensor*); };
enum { kDLCUDA = 2, kDLInt = 0, kDLUInt = 1, kDLFloat = 2 };

// The exported tensor holds a reference on the ARRAY, not on the var or the Ir: re-evaluating the var (commit_roots
// replaces a Binding root's array), dropping the var, or destroying the whole Ir only drops THEIR reference; the
// memory stays valid until the consumer calls the deleter, which needs neither the Ir nor its lock.
struct ExportCtx { Array* array; int64_t shape; };
void export_deleter(DLManagedTensor* t) {
  ExportCtx* c = (ExportCtx*)t->manager_ctx;
  release_array(c->array);  // the reference taken at export (exposed arrays are freed device-synchronously, runtime.cpp)
  drain_foreign_releases();
  delete c;
  delete t;
}
void import_release(void* p) {
  DLManagedTensor* t = (DLManagedTensor*)p;
  if (t->deleter) t->deleter(t);
}
}  // namespace

vkjit_status vkjit_var_to_dlpack(vkjit_ir* h, vkjit_var id, void** out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ir.is_buffer(id)) fail(VKJIT_ERR_INVALID, "to_dlpack: the var is not evaluated");
    const Var& v = ir.var(id);
    DLDataType dt{};
    switch (v.ty) {
      case VKJIT_TY_F32: dt = {kDLFloat, 32, 1}; break;
      case VKJIT_TY_I32: dt = {kDLInt, 32, 1}; break;
      case VKJIT_TY_U32: case VKJIT_TY_BOOL: dt = {kDLUInt, 32, 1}; break;  // Bool arrays hold 0/1 words (vartype.rs:45-64)
      default: fail(VKJIT_ERR_TYPE, "to_dlpack: scalar arrays only");
    }
    if (Backend::initialized()) Backend::get().sync();  // the consumer may use any stream
    auto* c = new ExportCtx{v.array, (int64_t)(v.array->bytes / 4)};
    auto* t = new DLManagedTensor();
    t->dl_tensor.data = v.array->ptr;
    v.array->exposed = true;  // the consumer may write it on its own streams from now on
    t->dl_tensor.device = {kDLCUDA, Backend::initialized() ? Backend::get().device : 0};
    t->dl_tensor.ndim = 1;
    t->dl_tensor.dtype = dt;
    t->dl_tensor.shape = &c->shape;
    t->dl_tensor.strides = nullptr;
    t->dl_tensor.byte_offset = 0;
    t->manager_ctx = c;
    t->deleter = export_deleter;
    retain_array(v.array);  // the tensor keeps the array alive until the consumer calls the deleter
    *out = t;
  });
}

vkjit_status vkjit_var_from_dlpack(vkjit_ir* h, void* managed, vkjit_var* out) {
  if (!managed) { g_last_error = "null DLManagedTensor"; return VKJIT_ERR_INVALID; }
  DLManagedTensor* t = (DLManagedTensor*)managed;
  const DLTensor& d = t->dl_tensor;
  vkjit_type ty = VKJIT_TY_VOID;
  const vkjit_status st = guard([&] {
    if (d.device.device_type != kDLCUDA) fail(VKJIT_ERR_INVALID, "from_dlpack: not CUDA device memory");
    if (Backend::initialized() && d.device.device_id != Backend::get().device) fail(VKJIT_ERR_INVALID, "from_dlpack: tensor lives on another GPU");
    if (d.dtype.lanes != 1 || d.dtype.bits != 32) fail(VKJIT_ERR_TYPE, "from_dlpack: 32-bit f32/i32/u32 elements only");
    ty = d.dtype.code == kDLFloat ? VKJIT_TY_F32 : d.dtype.code == kDLInt ? VKJIT_TY_I32 : d.dtype.code == kDLUInt ? VKJIT_TY_U32 : VKJIT_TY_VOID;
    if (ty == VKJIT_TY_VOID) fail(VKJIT_ERR_TYPE, "from_dlpack: 32-bit f32/i32/u32 elements only");
    if (d.ndim != 1) fail(VKJIT_ERR_INVALID, "from_dlpack: 1-D tensors only");
    if (d.strides && d.shape[0] > 1 && d.strides[0] != 1) fail(VKJIT_ERR_INVALID, "from_dlpack: the tensor is not contiguous");
  });
  if (st != VKJIT_OK) return st;  // the caller keeps ownership of the capsule
  return vkjit_array_wrap_device_owned(h, ty, (uint64_t)(uintptr_t)d.data + d.byte_offset, (size_t)d.shape[0], import_release, t, out);
}
void vkjit_dlpack_delete(void* managed) {
  DLManagedTensor* t = (DLManagedTensor*)managed;
  if (t && t->deleter) t->deleter(t);
}

vkjit_status vkjit_arange(vkjit_ir* h, vkjit_type ty, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.arange(ty, n, 0, false); }); }
vkjit_status vkjit_linspace(vkjit_ir* h, vkjit_type ty, vkjit_var a, vkjit_var b, size_t n, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.linspace(ty, a, b, n); });
}
vkjit_status vkjit_zeros(vkjit_ir* h, vkjit_type ty, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.zeros(ty); }); }
vkjit_status vkjit_ones(vkjit_ir* h, vkjit_type ty, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.ones(ty); }); }
vkjit_status vkjit_cast(vkjit_ir* h, vkjit_var s, vkjit_type ty, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.cast(s, ty); }); }
vkjit_status vkjit_bop(vkjit_ir* h, int32_t k, vkjit_var l, vkjit_var r, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.bop(k, l, r); }); }
vkjit_status vkjit_uop(vkjit_ir* h, int32_t k, vkjit_var s, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.uop(k, s); }); }
vkjit_status vkjit_bitcast(vkjit_ir* h, vkjit_var s, vkjit_type ty, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.bitcast(s, ty); }); }
vkjit_status vkjit_select(vkjit_ir* h, vkjit_var c, vkjit_var l, vkjit_var r, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.select(c, l, r); }); }
vkjit_status vkjit_struct_init(vkjit_ir* h, const vkjit_var* e, size_t n, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.struct_init(e, n); }); }
vkjit_status vkjit_getattr(vkjit_ir* h, vkjit_var s, size_t i, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.getattr(s, i); }); }
vkjit_status vkjit_setattr(vkjit_ir* h, vkjit_var d, vkjit_var s, size_t i, vkjit_var* out) { return with_ir(h, [&](Ir& ir) { *out = ir.setattr(d, s, i); }); }
vkjit_status vkjit_gather(vkjit_ir* h, vkjit_var s, vkjit_var i, int32_t ha, vkjit_var a, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.gather(s, i, ha != 0, a); });
}
vkjit_status vkjit_scatter(vkjit_ir* h, vkjit_var s, vkjit_var d, vkjit_var i, int32_t ha, vkjit_var a, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.scatter(OP_SCATTER, s, d, i, ha != 0, a); });
}
vkjit_status vkjit_scatter_add(vkjit_ir* h, vkjit_var s, vkjit_var d, vkjit_var i, int32_t ha, vkjit_var a, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) { *out = ir.scatter(OP_SCATTER_ADD, s, d, i, ha != 0, a); });
}

// ---- introspection / lifetime ---------------------------------------------------------------------
vkjit_status vkjit_var_type(vkjit_ir* h, vkjit_var id, vkjit_type* out) { return with_ir(h, [&](Ir& ir) { *out = ir.var(id).ty; }); }
vkjit_status vkjit_var_ref_count(vkjit_ir* h, vkjit_var id, uint32_t* out) {
  return with_ir(h, [&](Ir& ir) {
    if (id >= ir.vars.size()) fail(VKJIT_ERR_INVALID, "invalid VarId");
    *out = ir.vars[id].ref_count;  // readable for dead vars too (test.rs:205)
  });
}
vkjit_status vkjit_var_deps(vkjit_ir* h, vkjit_var id, vkjit_var* deps, size_t cap, size_t* out_ndeps, int32_t* out_has_se,
                            vkjit_var* out_se) {
  return with_ir(h, [&](Ir& ir) {
    const Var& v = ir.var(id);
    const bool leaf = v.op == OP_BINDING || v.op == OP_ARANGE;  // the inline dep slots hold num / base there
    const size_t nd = leaf ? 0 : v.ndeps;
    if (out_ndeps) *out_ndeps = nd;
    for (size_t k = 0; k < nd && k < cap; ++k) deps[k] = v.deps()[k];
    if (out_has_se) *out_has_se = (!leaf && v.has_se) ? 1 : 0;
    if (out_se) *out_se = (!leaf && v.has_se) ? v.side_effect : 0;
  });
}
vkjit_status vkjit_var_count(vkjit_ir* h, size_t* out) { return with_ir(h, [&](Ir& ir) { *out = ir.vars.size(); }); }
vkjit_status vkjit_array_count(vkjit_ir* h, size_t* out) { return with_ir(h, [&](Ir& ir) { *out = ir.n_arrays; }); }
vkjit_status vkjit_is_buffer(vkjit_ir* h, vkjit_var id, int32_t* out) { return with_ir(h, [&](Ir& ir) { ir.var(id); *out = ir.is_buffer(id); }); }
vkjit_status vkjit_var_size(vkjit_ir* h, vkjit_var id, size_t* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ir.is_buffer(id)) fail(VKJIT_ERR_INVALID, "var is not a buffer");
    *out = ir.var(id).array->bytes / 4;
  });
}
vkjit_status vkjit_var_device_ptr(vkjit_ir* h, vkjit_var id, uint64_t* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ir.is_buffer(id)) fail(VKJIT_ERR_INVALID, "var is not a buffer");
    ir.var(id).array->exposed = true;  // whoever holds the pointer may write the memory (CUDA Array Interface consumers)
    *out = (uint64_t)(uintptr_t)ir.var(id).array->ptr;
  });
}
vkjit_status vkjit_inc_ref(vkjit_ir* h, vkjit_var id) { return with_ir(h, [&](Ir& ir) { ir.inc_ref(id); }); }
vkjit_status vkjit_dec_ref(vkjit_ir* h, vkjit_var id) { return with_ir(h, [&](Ir& ir) { ir.dec_ref(id); }); }
vkjit_status vkjit_ir_repr(vkjit_ir* h, char* buf, size_t cap, size_t* out_len) {
  return with_ir(h, [&](Ir& ir) { copy_out(ir.repr(), buf, cap, out_len); });
}
vkjit_status vkjit_var_repr(vkjit_ir* h, vkjit_var id, char* buf, size_t cap, size_t* out_len) {
  return with_ir(h, [&](Ir& ir) { copy_out(ir.is_buffer(id) ? buffer_str(ir, id) : ir.var_debug(id), buf, cap, out_len); });
}

// ---- execute -----------------------------------------------------------------------------------------
vkjit_status vkjit_schedule(vkjit_ir* h, const vkjit_var* ids, size_t n) { return with_ir(h, [&](Ir& ir) { ir.do_schedule(ids, n); }); }
vkjit_status vkjit_eval(vkjit_ir* h, const vkjit_var* ids, size_t n) { return with_ir(h, [&](Ir& ir) { eval(ir, ids, n); }); }
// Ir::as_slice<T> (internal.rs:443-449)
vkjit_status vkjit_read(vkjit_ir* h, vkjit_var id, vkjit_type ty, void* dst, size_t bytes) {
  return with_ir(h, [&](Ir& ir) {
    if (!ir.is_buffer(id)) fail(VKJIT_ERR_INVALID, "as_slice on a var that is not a buffer (internal.rs:446)");
    const Var& v = ir.var(id);
    if (v.ty != ty) fail(VKJIT_ERR_TYPE, "as_slice type mismatch (internal.rs:447)");
    Backend::get().d2h(dst, v.array->ptr, std::min(bytes, v.array->bytes));
  });
}

// ---- runtime primitives ----------------------------------------------------------------------------------
// identity element of a reduction, as a 4-byte pattern NCCL's min/max/sum treat as neutral
static uint32_t reduce_identity(int red, TypeId ty) {
  if (red == VKJIT_RED_SUM) return 0u;
  if (ty == VKJIT_TY_F32) return red == VKJIT_RED_MIN ? 0x7F800000u : 0xFF800000u;  // +inf / -inf
  if (ty == VKJIT_TY_I32) return red == VKJIT_RED_MIN ? 0x7FFFFFFFu : 0x80000000u;
  return red == VKJIT_RED_MIN ? 0xFFFFFFFFu : 0u;
}

// $VKJIT_REDUCE_OVERLAP=0 switches the programmatic-dependent-launch overlap of consecutive reductions off (A/B)
static bool reduce_overlap_enabled() {
  static const bool on = [] { const char* e = getenv("VKJIT_REDUCE_OVERLAP"); return !(e && e[0] == '0'); }();
  return on;
}

vkjit_status vkjit_reduce(vkjit_ir* h, int32_t red, vkjit_var id, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    const TypeId ty = ir.var(id).ty;
    if (!ty_is_num(ty)) fail(VKJIT_ERR_TYPE, "reduce needs U32/I32/F32");
    if (red < VKJIT_RED_SUM || red > VKJIT_RED_MAX) fail(VKJIT_ERR_INVALID, "unknown reduction");
    Backend& be = Backend::get();
    commit_if_side_effects(ir, &id, 1);
    const bool sharded = ir.var(id).sharded;
    const bool combine = sharded && dist::active() && dist::world() > 1;
    // a rank whose shard of a tiny array is empty still has to take part in the collective
    bool empty;
    if (ir.is_buffer(id)) empty = ir.var(id).array->bytes == 0;
    else {
      Program p;
      std::vector<VarId> roots{id};
      build_program(ir, roots, true, p);
      empty = p.n == 0;
    }
    if (empty && !combine) fail(VKJIT_ERR_SIZE, "reduce of an empty array");
    const bool p2p = combine && dist::p2p_enabled();
    prims::Mailbox mb;
    if (p2p) mb = dist::next_mailbox();
    Array* o = nullptr;
    bool counted = false;
    try {
      if (empty) {
        o = be.new_array(4);
        prims::fill_u32((uint32_t*)o->ptr, reduce_identity(red, ty), 1, be.enqueue_stream());
        if (p2p) prims::p2p_allreduce(red, ty, o->ptr, mb, be.enqueue_stream());
      } else if (!ir.is_buffer(id) || misaligned(ir, id)) {
        // unevaluated operand (or a misaligned foreign view): ONE generated kernel evaluates the trace and reduces
        // it; the operand is not materialised and stays as it is
        o = eval_reduce(ir, id, red);
        if (p2p) prims::p2p_allreduce(red, ty, o->ptr, mb, be.enqueue_stream());
      } else {
        o = be.new_array(4);
        const Var& v = ir.var(id);
        // Back-to-back reductions overlap (prims.cu: reduce_kernel, programmatic dependent launch): the streaming
        // phase of this one may run under the tail of the previous one if (a) that previous kernel is a reduction
        // of this chain — nothing else went onto the stream since, (b) the input is none of the results the chain
        // still has in flight, and (c) nobody outside this backend can be writing the input (foreign views,
        // exported pointers).  Otherwise the kernel waits first: plain stream order.
        Backend::ReduceChain& ch = be.reduce_chain;
        std::lock_guard<std::mutex> chain_lock(ch.mu);
        const void* in = v.array->ptr;
        // (d) the input is large enough for the overlap to matter (>= 2^16 lanes): a reduction's own 1-element result
        // is never an overlapped input
        bool overlap = reduce_overlap_enabled() && ch.sig == Backend::counters().stream_ops && v.array->owned && !v.array->exposed &&
                       v.array->bytes >= (size_t(1) << 18) && ch.outs.size() < 32;
        for (size_t i = 0; overlap && i < ch.outs.size(); ++i) overlap = ch.outs[i] != in;
        if (!overlap) ch.outs.clear();
        // p2p: the last CTA of the reduction exchanges the per-GPU partial over NVLink peer memory
        prims::reduce(red, ty, in, v.array->bytes / 4, o->ptr, be.scratch, be.sm_count, be.enqueue_stream(), p2p ? &mb : nullptr,
                      overlap ? 0u : prims::kReduceWaitFirst);
        Backend::counters().note_prim();
        ch.outs.push_back(o->ptr);
        ch.sig = (combine && !p2p) ? ~0ull : (uint64_t)Backend::counters().stream_ops;  // an NCCL kernel follows: chain ends
        counted = true;
      }
      if (!counted) Backend::counters().note_prim();
      if (combine && !p2p) dist::allreduce(o->ptr, ty, red, 1);  // NCCL: per-GPU partial -> replicated result
    } catch (...) { release_array(o); throw; }
    *out = ir.binding(ty, o, false);
  });
}

// Exclusive scan over ranks of one u32 per GPU: out[0] = sum of *mine over the lower ranks and, with `total`,
// out[1] = sum over all ranks.  Through the NVLink peer mailboxes, or as one-hot vector -> ncclAllReduce -> prefix.
// `vec`: `world` words of device scratch (NCCL path).  mine may alias out.
void rank_exscan(Backend& be, const uint32_t* mine, uint32_t* out, bool total, uint32_t* vec) {
  const int world = dist::world(), rank = dist::rank();
  if (dist::p2p_enabled()) {
    if (total) prims::p2p_exscan_total_u32(mine, out, dist::next_mailbox(), be.enqueue_stream());
    else prims::p2p_exscan_u32(mine, out, dist::next_mailbox(), be.enqueue_stream());
  } else {
    prims::one_hot_u32(mine, rank, world, vec, be.enqueue_stream());
    dist::allreduce(vec, VKJIT_TY_U32, VKJIT_RED_SUM, (size_t)world);
    prims::prefix_of_rank_u32(vec, rank, out, be.enqueue_stream());
    if (total) prims::prefix_of_rank_u32(vec, world, out + 1, be.enqueue_stream());
  }
  Backend::counters().note_prim();
}

// Operand of an eager primitive as a 16-byte aligned device array: the var's own array, or — for an unevaluated
// var or a misaligned foreign view — a temporary evaluated on the fly.  The var itself is never changed: like
// reduce, the primitives do not turn an unevaluated operand into a buffer (oracle.cpp ditto).
struct Operand {
  const uint32_t* ptr = nullptr;
  size_t n = 0;
  Array* temp = nullptr;
  Operand() = default;
  Operand(const Operand&) = delete;
  Operand& operator=(const Operand&) = delete;
  ~Operand() { if (temp) release_array(temp); }  // stream-ordered: kernels already enqueued still see the memory
  void bind(Ir& ir, VarId id) {
    if (ir.is_buffer(id) && !misaligned(ir, id)) {
      const Array* a = ir.var(id).array;
      ptr = (const uint32_t*)a->ptr; n = a->bytes / 4;
    } else {
      temp = eval_temp(ir, id);
      ptr = (const uint32_t*)temp->ptr; n = temp->bytes / 4;
    }
  }
};

vkjit_status vkjit_prefix_sum(vkjit_ir* h, vkjit_var id, int32_t exclusive, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    const TypeId ty = ir.var(id).ty;
    if (ty != VKJIT_TY_U32 && ty != VKJIT_TY_I32) fail(VKJIT_ERR_TYPE, "prefix_sum needs U32/I32");
    Backend& be = Backend::get();
    commit_if_side_effects(ir, &id, 1);
    const bool sharded = ir.var(id).sharded && dist::active() && dist::world() > 1;
    // a rank whose shard is empty cannot evaluate it, but still has to take part in the exchange
    bool empty;
    if (ir.is_buffer(id)) empty = ir.var(id).array->bytes == 0;
    else {
      Program p;
      std::vector<VarId> roots{id};
      build_program(ir, roots, true, p);
      empty = p.n == 0;
    }
    const bool direct = !empty && ir.is_buffer(id) && !misaligned(ir, id);  // the hand-written kernel reads the var's array
    Operand in;
    if (direct) in.bind(ir, id);
    Array* o = nullptr;
    Array* total = nullptr;  // sharded, unevaluated operand: local total from the fused trace -> reduce kernel
    void* tmp = nullptr;     // sharded: [0] local total, [1] offset of this rank, [2..2+world) totals (NCCL path)
    const size_t tmp_bytes = (size_t)(2 + (sharded ? dist::world() : 0)) * 4;
    try {
      const uint32_t* initial = nullptr;
      if (sharded) {
        // Sharded scan: local total -> exchange of the per-rank totals -> single-pass scan whose
        // tile 0 starts from the sum of the lower ranks' totals.  12 B/lane per GPU, one small exchange.
        tmp = be.alloc(tmp_bytes);
        uint32_t* w = (uint32_t*)tmp;
        const uint32_t* tot = w;
        if (empty) prims::fill_u32(w, 0u, 1, be.enqueue_stream());
        else if (direct) prims::reduce(VKJIT_RED_SUM, VKJIT_TY_U32, in.ptr, in.n, w, be.scratch, be.sm_count, be.enqueue_stream());
        else { total = eval_reduce(ir, id, VKJIT_RED_SUM); tot = (const uint32_t*)total->ptr; }
        rank_exscan(be, tot, w + 1, false, w + 2);
        initial = w + 1;
        Backend::counters().note_prim();
      }
      bool done = false;
      if (!empty && !direct) {  // ONE generated kernel evaluates the trace and scans it: the addends never reach memory
        std::vector<VarId> roots{id};
        uint64_t n = 0;
        done = eval_scan(ir, exclusive ? SCAN_EXCLUSIVE : SCAN_INCLUSIVE, roots, initial, nullptr, &o, &n);
        if (!done) in.bind(ir, id);
      }
      if (!done) {
        be.ensure_scan_scratch(in.n);
        o = be.new_array(in.n * 4);
        prims::prefix_sum(in.ptr, (uint32_t*)o->ptr, in.n, exclusive != 0, be.scratch, be.sm_count, be.enqueue_stream(), initial);
      }
      Backend::counters().note_prim();
    } catch (...) {
      if (tmp) be.free_async(tmp, tmp_bytes);
      release_array(total); release_array(o);
      throw;
    }
    if (tmp) be.free_async(tmp, tmp_bytes);
    release_array(total);
    *out = ir.binding(ty, o, sharded);
  });
}

static void do_compress(Ir& ir, bool with_values, VarId values, VarId mask, vkjit_var* out, size_t* count) {
  if (ir.var(mask).ty != VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "compress mask must be Bool");
  TypeId oty = VKJIT_TY_U32;
  if (with_values) {
    oty = ir.var(values).ty;
    if (!ty_is_scalar(oty)) fail(VKJIT_ERR_TYPE, "compress values must be scalar");
  }
  Backend& be = Backend::get();
  {
    const VarId ops[2] = {mask, values};
    commit_if_side_effects(ir, ops, with_values ? 2 : 1);
  }
  // Sharded compress: every rank compacts its own shard; the result is a ragged sharded array —
  // rank r holds the global elements [offset_r, offset_r + count_r), `count` is the GLOBAL number of selected lanes,
  // index results are global lane numbers.  Two small exchanges: exscan of the shard sizes (index base) and
  // exscan + total of the counts.
  const bool sharded = ir.var(mask).sharded && dist::active() && dist::world() > 1;
  if (sharded && with_values && !ir.var(values).sharded) fail(VKJIT_ERR_SIZE, "compress: sharded mask with unsharded values");
  size_t n_local = 0;
  if (sharded) {  // a rank whose shard is empty cannot evaluate it, but still has to take part in the exchanges
    if (ir.is_buffer(mask)) n_local = ir.var(mask).array->bytes / 4;
    else {
      Program p;
      std::vector<VarId> roots{mask};
      build_program(ir, roots, true, p);
      n_local = p.n;
    }
  }
  const bool direct = ir.is_buffer(mask) && !misaligned(ir, mask) && (!with_values || (ir.is_buffer(values) && !misaligned(ir, values)));
  Array* o = nullptr;
  // device words: [0] local count, [1] offset of this rank, [2] global count, [3] index base, [4..4+world) NCCL scratch
  const size_t cnt_bytes = (size_t)(4 + (sharded ? dist::world() : 0)) * 4;
  uint32_t* w = (uint32_t*)be.alloc(cnt_bytes);
  uint32_t host[3] = {0, 0, 0};
  try {
    const uint32_t* index_base = nullptr;
    if (sharded && !with_values) {
      prims::fill_u32(w + 3, (uint32_t)n_local, 1, be.enqueue_stream());
      rank_exscan(be, w + 3, w + 3, false, w + 4);
      index_base = w + 3;
    }
    uint64_t n = 0;
    bool done = false;
    if (sharded && n_local == 0) {
      o = be.new_array(0);
      done = true;
    } else if (!direct) {  // fused: the mask (and the values) are computed inside the compaction kernel
      std::vector<VarId> roots{mask};
      if (with_values) roots.push_back(values);
      try {
        done = eval_scan(ir, with_values ? SCAN_COMPRESS_VALUE : SCAN_COMPRESS_INDEX, roots, nullptr, w, &o, &n, index_base);
      } catch (const Error& e) {
        if (e.code == VKJIT_ERR_SIZE && with_values) fail(VKJIT_ERR_SIZE, "compress: values and mask sizes differ");
        throw;
      }
    }
    if (!done) {
      Operand m, v;
      m.bind(ir, mask);
      if (with_values) {
        v.bind(ir, values);
        if (v.n != m.n) fail(VKJIT_ERR_SIZE, "compress: values and mask sizes differ");
      }
      n = m.n;
      be.ensure_scan_scratch(n);
      o = be.new_array(n * 4);  // worst case; logical size is trimmed to the count below
      if (n) prims::compress(m.ptr, with_values ? v.ptr : nullptr, (uint32_t*)o->ptr, w, n, be.scratch, be.sm_count, be.enqueue_stream(), index_base);
    }
    if (n) Backend::counters().note_prim();
    else if (sharded) prims::fill_u32(w, 0u, 1, be.enqueue_stream());
    if (sharded) rank_exscan(be, w, w + 1, true, w + 4);
    // the size of the result is data dependent: one small readback
    if (n || sharded) be.d2h(host, w, sharded ? 12 : 4);
  } catch (...) { be.free_async(w, cnt_bytes); release_array(o); throw; }
  be.free_async(w, cnt_bytes);
  o->bytes = (size_t)host[0] * 4;
  *count = sharded ? host[2] : host[0];
  *out = ir.binding(oty, o, sharded);
  if (sharded) ir.var(*out).base = host[1];
}

vkjit_status vkjit_compress(vkjit_ir* h, vkjit_var mask, vkjit_var* out, size_t* count) {
  return with_ir(h, [&](Ir& ir) { do_compress(ir, false, 0, mask, out, count); });
}
vkjit_status vkjit_compress_values(vkjit_ir* h, vkjit_var values, vkjit_var mask, vkjit_var* out, size_t* count) {
  return with_ir(h, [&](Ir& ir) { do_compress(ir, true, values, mask, out, count); });
}

// ---- multi-GPU ------------------------------------------------------------------------------------------------
vkjit_status vkjit_dist_unique_id(void* out) { return guard([&] { dist::unique_id(out); }); }
vkjit_status vkjit_dist_init(int32_t rank, int32_t world, const void* id) { return guard([&] { dist::init(rank, world, id); }); }
vkjit_status vkjit_dist_mailbox_handle(void* out64) { return guard([&] { dist::mailbox_handle(out64); }); }
vkjit_status vkjit_dist_mailbox_open(const void* handles, int32_t world) { return guard([&] { dist::mailbox_open(handles, world); }); }
vkjit_status vkjit_dist_init_env(void) { return guard([&] { dist::init_env(); }); }
vkjit_status vkjit_debug_rendezvous(int32_t rank, int32_t world, const char* addr, int32_t port, const void* blob64, void* root128,
                                    void* out_all, double timeout_s) {
  return guard([&] { dist::rendezvous(rank, world, addr, port, blob64, 64, root128, 128, out_all, timeout_s); });
}
vkjit_status vkjit_dist_set_p2p(int32_t on) { return guard([&] { dist::set_p2p(on != 0); }); }
vkjit_status vkjit_dist_shutdown(void) { return guard([&] { dist::shutdown(); }); }
vkjit_status vkjit_dist_info(int32_t* rank, int32_t* world) {
  return guard([&] { *rank = dist::rank(); *world = dist::world(); });
}
vkjit_status vkjit_shard_range(size_t n, int32_t rank, int32_t world, size_t* lo, size_t* hi) {
  return guard([&] { dist::shard_range(n, rank, world, *lo, *hi); });
}
vkjit_status vkjit_arange_sharded(vkjit_ir* h, vkjit_type ty, size_t n, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    size_t lo, hi;
    dist::shard_range(n, dist::rank(), dist::world(), lo, hi);
    *out = ir.arange(ty, hi - lo, lo, true);
  });
}
vkjit_status vkjit_array_sharded(vkjit_ir* h, vkjit_type ty, const void* data, size_t n, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_scalar(ty)) fail(VKJIT_ERR_TYPE, "array of a non-scalar type");
    size_t lo, hi;
    dist::shard_range(n, dist::rank(), dist::world(), lo, hi);
    *out = upload(ir, ty, (const char*)data + lo * 4, hi - lo, true);
    ir.var(*out).base = lo;
  });
}
vkjit_status vkjit_array_shard_local(vkjit_ir* h, vkjit_type ty, const void* data, size_t n_local, vkjit_var* out) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_scalar(ty)) fail(VKJIT_ERR_TYPE, "array of a non-scalar type");
    *out = upload(ir, ty, data, n_local, true);
  });
}
vkjit_status vkjit_var_is_sharded(vkjit_ir* h, vkjit_var id, int32_t* out) { return with_ir(h, [&](Ir& ir) { *out = ir.var(id).sharded; }); }
vkjit_status vkjit_var_shard_base(vkjit_ir* h, vkjit_var id, uint64_t* out) {
  return with_ir(h, [&](Ir& ir) {
    const Var& v = ir.var(id);
    *out = (v.op == OP_BINDING || v.op == OP_ARANGE) ? v.base : 0;  // the field shares storage with the dependency ids
  });
}

// ---- counters ----------------------------------------------------------------------------------------------------
vkjit_status vkjit_stats(vkjit_stats_t* out) {
  return guard([&] {
    Counters& c = Backend::counters();
    out->cache_hits = c.cache_hits; out->cache_misses = c.cache_misses;
    out->trace_launches = c.trace_launches; out->prim_launches = c.prim_launches;
    out->last_compile_ns = c.last_compile_ns; out->last_eval_ns = c.last_eval_ns;
    out->bytes_h2d = c.bytes_h2d; out->bytes_d2h = c.bytes_d2h;
    out->pool_bytes_live = c.pool_bytes_live; out->collectives = c.collectives; out->disk_hits = c.disk_hits;
  });
}
vkjit_status vkjit_stats_reset(void) {
  return guard([&] {
    Counters& c = Backend::counters();
    c.cache_hits = 0; c.cache_misses = 0; c.trace_launches = 0; c.prim_launches = 0;
    c.last_compile_ns = 0; c.last_eval_ns = 0; c.bytes_h2d = 0; c.bytes_d2h = 0; c.collectives = 0; c.disk_hits = 0;
  });
}
vkjit_status vkjit_cache_clear(void) { return guard([&] { Backend::get().clear_cache(); }); }

vkjit_status vkjit_debug_codegen(vkjit_ir* h, const vkjit_var* ids, size_t n, int32_t compile, char* buf, size_t cap,
                                 size_t* out_len, size_t* out_cubin) {
  return with_ir(h, [&](Ir& ir) {
    std::vector<VarId> sched(ir.schedule);
    for (size_t i = 0; i < n; ++i) {
      ir.var(ids[i]);
      if (std::find(sched.begin(), sched.end(), ids[i]) == sched.end()) sched.push_back(ids[i]);
    }
    Program p;
    build_program(ir, sched, true, p, -1, (compile & 8) ? 3 : (compile & 4) ? 2 : (compile & 2) ? 1 : 0);  // compile bit 1: privatised scatter_add variant; bit 2: 2-CTA cluster variant; bit 3: bin-range passes
    const std::string src = generate_cuda(ir, p);
    if (out_cubin) *out_cubin = 0;
    if (compile & 1) {
      std::vector<char> cubin;
      std::string log;
      if (!nvrtc_compile(src, cubin, log)) fail(VKJIT_ERR_COMPILE, "NVRTC rejected the generated kernel:\n" + log + "\n--- source ---\n" + src);
      if (out_cubin) *out_cubin = cubin.size();
    }
    copy_out(src, buf, cap, out_len);
  });
}

vkjit_status vkjit_debug_walk_ns(vkjit_ir* h, const vkjit_var* ids, size_t n, uint32_t reps, uint64_t* out_ns, uint32_t* out_nodes) {
  return with_ir(h, [&](Ir& ir) {
    std::vector<VarId> sched(ids, ids + n);
    static thread_local Program p;  // warm buffers, as in eval
    if (reps == 0) {  // one walk of a trace in whatever cache state its construction left it: what a fresh eval pays
      const uint64_t t0 = now_ns();
      build_program(ir, sched, true, p);
      *out_ns = now_ns() - t0;
      *out_nodes = (uint32_t)p.order.size();
      return;
    }
    build_program(ir, sched, true, p);
    const uint64_t t0 = now_ns();
    for (uint32_t i = 0; i < reps; ++i) build_program(ir, sched, true, p);
    *out_ns = (now_ns() - t0) / reps;
    *out_nodes = (uint32_t)p.order.size();
  });
}

vkjit_status vkjit_debug_reduce_trace(uint64_t* out, size_t cap_words, size_t* out_launches) {
  return guard([&] {
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "64-bit stamps");
    *out_launches = prims::reduce_trace_dump((unsigned long long*)out, cap_words, Backend::get().wait_stream());
  });
}

vkjit_status vkjit_debug_eval_bookkeeping(vkjit_ir* h, const vkjit_var* ids, size_t n) {
  return with_ir(h, [&](Ir& ir) {
    ir.do_schedule(ids, n);
    if (ir.schedule.empty()) return;
    try {
      Program p;
      build_program(ir, ir.schedule, true, p);
      std::vector<Array*> outs;
      for (size_t r = 0; r < p.roots.size(); ++r) {
        Array* a = new Array();  // a foreign view of nothing: the right size for later walks, no memory behind it
        a->owned = false;
        a->bytes = a->capacity = (size_t)p.n * 4;
        outs.push_back(a);
      }
      ir.commit_roots(ir.schedule, outs, p.order);
    } catch (...) {
      ir.clear_schedule();
      throw;
    }
    ir.clear_schedule();
  });
}

vkjit_status vkjit_debug_codegen_reduce(vkjit_ir* h, vkjit_var id, int32_t red, int32_t compile, char* buf, size_t cap,
                                        size_t* out_len, size_t* out_cubin) {
  return with_ir(h, [&](Ir& ir) {
    if (!ty_is_num(ir.var(id).ty)) fail(VKJIT_ERR_TYPE, "reduce needs U32/I32/F32");
    if (red < VKJIT_RED_SUM || red > VKJIT_RED_MAX) fail(VKJIT_ERR_INVALID, "unknown reduction");
    Program p;
    std::vector<VarId> roots{id};
    build_program(ir, roots, true, p, red);
    const std::string src = generate_cuda(ir, p);
    if (out_cubin) *out_cubin = 0;
    if (compile) {
      std::vector<char> cubin;
      std::string log;
      if (!nvrtc_compile(src, cubin, log)) fail(VKJIT_ERR_COMPILE, "NVRTC rejected the generated kernel:\n" + log + "\n--- source ---\n" + src);
      if (out_cubin) *out_cubin = cubin.size();
    }
    copy_out(src, buf, cap, out_len);
  });
}

vkjit_status vkjit_debug_codegen_scan(vkjit_ir* h, const vkjit_var* ids, size_t n, int32_t mode, int32_t compile, char* buf,
                                      size_t cap, size_t* out_len, size_t* out_cubin) {
  return with_ir(h, [&](Ir& ir) {
    if (mode < SCAN_EXCLUSIVE || mode > SCAN_COMPRESS_VALUE) fail(VKJIT_ERR_INVALID, "unknown scan mode");
    if (n != (mode == SCAN_COMPRESS_VALUE ? 2u : 1u)) fail(VKJIT_ERR_INVALID, "scan modes 0-2 take one var, mode 3 takes {mask, values}");
    Program p;
    std::vector<VarId> roots(ids, ids + n);
    build_program(ir, roots, true, p, -1, false, mode);
    const std::string src = generate_cuda(ir, p);
    if (out_cubin) *out_cubin = 0;
    if (compile) {
      std::vector<char> cubin;
      std::string log;
      if (!nvrtc_compile(src, cubin, log)) fail(VKJIT_ERR_COMPILE, "NVRTC rejected the generated kernel:\n" + log + "\n--- source ---\n" + src);
      if (out_cubin) *out_cubin = cubin.size();
    }
    copy_out(src, buf, cap, out_len);
  });
}

}  // extern "C"
