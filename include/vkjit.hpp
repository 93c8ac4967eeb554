// vkjit.hpp — header-only C++17 front-end over the C ABI (include/vkjit_b200.h).
//
// Host-side mirror of the reference's compiled front-end crate `vkjit-rust`
// (libs/vkjit-rust/src/{lib,types,functions}.rs): one process-wide `Ir` (lib.rs:9-15), a `Var` that
// owns exactly one reference count of its VarId (Clone = inc_ref_count, Drop = dec_ref_count,
// types.rs:128-140), arithmetic operators for anything convertible into a `Var` (types.rs:42-73),
// the named comparisons (types.rs:75-88), `getattr / setattr / then_else / scatter / scatter_with /
// get / to_vec` (types.rs:142-197) and the free functions `zeros, arange, linspace, select, gather,
// gather_with, repr_ir, eval, schedule` (functions.rs:5-82; `eval!` / `schedule!` are variadic
// functions here).  The reference panics on misuse; this header throws `vkjit::Error` carrying
// the ABI status and message.  Extensions the reference lacks (reductions, prefix sum, compress,
// bit / unary ops) are at the end.  There is no CPU path: without a B200 traces can be built and
// printed, uploads and `eval` throw with VKJIT_ERR_NO_DEVICE.
//
//   #include "vkjit.hpp"
//   using namespace vkjit;
//   Var x = arange(U32, 10);           // src/main.rs of the reference
//   eval(x);
//   std::cout << x << "\n";             // Var("[0, 1, 2, 3, 4, 5, 6, 7, 8, 9]")
#pragma once
#include <cstdint>
#include <initializer_list>
#include <ostream>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>
#include <vector>

#include "vkjit_b200.h"

namespace vkjit {

struct Error : std::runtime_error {
  vkjit_status status;
  Error(vkjit_status st, const std::string& msg) : std::runtime_error("[status " + std::to_string(st) + "] " + msg), status(st) {}
};

namespace detail {
inline void check(vkjit_status st) {
  if (st != VKJIT_OK) {
    const char* m = vkjit_last_error();
    throw Error(st, m ? m : "");
  }
}

// `lazy_static! { pub static ref IR: Mutex<Ir> }` (lib.rs:9-15).  Ir::new() creates the backend
// (internal.rs:168-182); the device is bound here on first use if the process has not done so.  The
// handle lives until process exit, like the reference's static.
inline vkjit_ir* ir() {
  static vkjit_ir* handle = [] {
    if (!vkjit_is_initialized()) {
      const vkjit_status st = vkjit_init(-1);
      if (st != VKJIT_OK && st != VKJIT_ERR_NO_DEVICE) check(st);  // no device: construction still works, eval throws
    }
    vkjit_ir* h = nullptr;
    check(vkjit_ir_create(&h));
    return h;
  }();
  return handle;
}

template <class T> struct type_of;
template <> struct type_of<float> { static constexpr vkjit_type value = VKJIT_TY_F32; };
template <> struct type_of<int32_t> { static constexpr vkjit_type value = VKJIT_TY_I32; };
template <> struct type_of<uint32_t> { static constexpr vkjit_type value = VKJIT_TY_U32; };

template <class F> std::string text_of(F&& call) {
  size_t len = 0;
  check(call(nullptr, 0, &len));
  std::string s(len + 1, '\0');
  check(call(&s[0], s.size(), &len));
  s.resize(len);
  return s;
}
}  // namespace detail

// VarType (vartype.rs:24-33): scalar codes of the ABI; struct types are interned per Ir
using VarType = vkjit_type;
constexpr VarType Void = VKJIT_TY_VOID, Bool = VKJIT_TY_BOOL, U32 = VKJIT_TY_U32, I32 = VKJIT_TY_I32, F32 = VKJIT_TY_F32;
inline VarType Struct(std::initializer_list<VarType> members) {  // VarType::Struct(vec![...])
  std::vector<VarType> m(members);
  VarType t = 0;
  detail::check(vkjit_type_struct(detail::ir(), m.data(), m.size(), &t));
  return t;
}

class Var {
 public:
  // From<VarId> (types.rs:107-111): adopts an id that already carries one reference
  static Var from_id(vkjit_var id) { return Var(id, Adopt{}); }

  // From<f32 | i32 | u32 | bool> (types.rs:90-93)
  Var(float v) { detail::check(vkjit_const_f32(detail::ir(), v, &id_)); }
  Var(double v) : Var(static_cast<float>(v)) {}  // a Rust float literal is inferred as f32; a C++ one is a double
  Var(int32_t v) { detail::check(vkjit_const_i32(detail::ir(), v, &id_)); }
  Var(uint32_t v) { detail::check(vkjit_const_u32(detail::ir(), v, &id_)); }
  Var(bool v) { detail::check(vkjit_const_bool(detail::ir(), v ? 1 : 0, &id_)); }
  // From<&[T]> / From<Vec<T>> for f32 | i32 | u32 (types.rs:95-97): upload
  Var(const std::vector<float>& v) { detail::check(vkjit_array_f32(detail::ir(), v.data(), v.size(), &id_)); }
  Var(const std::vector<int32_t>& v) { detail::check(vkjit_array_i32(detail::ir(), v.data(), v.size(), &id_)); }
  Var(const std::vector<uint32_t>& v) { detail::check(vkjit_array_u32(detail::ir(), v.data(), v.size(), &id_)); }
  Var(const float* p, size_t n) { detail::check(vkjit_array_f32(detail::ir(), p, n, &id_)); }
  Var(const int32_t* p, size_t n) { detail::check(vkjit_array_i32(detail::ir(), p, n, &id_)); }
  Var(const uint32_t* p, size_t n) { detail::check(vkjit_array_u32(detail::ir(), p, n, &id_)); }
  // From<&[Var]> (types.rs:99-105): a struct-typed var from its members
  static Var structure(const std::vector<Var>& members) {
    std::vector<vkjit_var> ids;
    for (const Var& m : members) ids.push_back(m.id());
    vkjit_var out = 0;
    detail::check(vkjit_struct_init(detail::ir(), ids.data(), ids.size(), &out));
    return from_id(out);
  }

  Var(const Var& o) : id_(o.id_) { detail::check(vkjit_inc_ref(detail::ir(), id_)); }  // Clone (types.rs:128-133)
  Var(Var&& o) noexcept : id_(o.id_), live_(o.live_) { o.live_ = false; }
  Var& operator=(Var o) noexcept {  // by value: the old variable is released when `o` dies
    std::swap(id_, o.id_);
    std::swap(live_, o.live_);
    return *this;
  }
  ~Var() {  // Drop (types.rs:135-140)
    if (live_) vkjit_dec_ref(detail::ir(), id_);
  }

  vkjit_var id() const { return id_; }
  VarType ty() const {
    VarType t = 0;
    detail::check(vkjit_var_type(detail::ir(), id_, &t));
    return t;
  }
  bool is_buffer() const {
    int32_t b = 0;
    detail::check(vkjit_is_buffer(detail::ir(), id_, &b));
    return b != 0;
  }
  size_t size() const {  // elements of an evaluated var
    size_t n = 0;
    detail::check(vkjit_var_size(detail::ir(), id_, &n));
    return n;
  }

  Var getattr(size_t idx) const {  // types.rs:149-151
    vkjit_var out = 0;
    detail::check(vkjit_getattr(detail::ir(), id_, idx, &out));
    return from_id(out);
  }
  void setattr(const Var& v, size_t idx) {  // types.rs:152-159: `*self = ret`
    vkjit_var out = 0;
    detail::check(vkjit_setattr(detail::ir(), id_, v.id(), idx, &out));
    *this = from_id(out);
  }
  Var then_else(const Var& then, const Var& other) const {  // types.rs:160-168
    vkjit_var out = 0;
    detail::check(vkjit_select(detail::ir(), id_, then.id(), other.id(), &out));
    return from_id(out);
  }
  // types.rs:169-189: the value of the scatter var is its source; reading the TARGET shows the effect.
  void scatter(const Var& to, const Var& idx) { scatter_impl(to, idx, nullptr); }
  void scatter_with(const Var& to, const Var& idx, const Var& condition) { scatter_impl(to, idx, &condition); }
  Var get(const Var& idx) const;  // types.rs:190-192 = gather(self.clone(), idx)

  // types.rs:193-196 (Ir::as_slice::<T>): the var must be evaluated and of type T
  template <class T> std::vector<T> to_vec() const {
    std::vector<T> out(size());
    detail::check(vkjit_read(detail::ir(), id_, detail::type_of<T>::value, out.data(), out.size() * sizeof(T)));
    return out;
  }

  // named_bop! (types.rs:75-88)
  Var lt(const Var& r) const { return bop(VKJIT_BOP_LT, r); }
  Var gt(const Var& r) const { return bop(VKJIT_BOP_GT, r); }
  Var eq(const Var& r) const { return bop(VKJIT_BOP_EQ, r); }
  Var leq(const Var& r) const { return bop(VKJIT_BOP_LEQ, r); }
  Var geq(const Var& r) const { return bop(VKJIT_BOP_GEQ, r); }
  Var neq(const Var& r) const { return bop(VKJIT_BOP_NEQ, r); }

  // bop!(Add | Sub | Mul | Div) + the *Assign forms (types.rs:42-73)
  Var& operator+=(const Var& r) { return *this = bop(VKJIT_BOP_ADD, r); }
  Var& operator-=(const Var& r) { return *this = bop(VKJIT_BOP_SUB, r); }
  Var& operator*=(const Var& r) { return *this = bop(VKJIT_BOP_MUL, r); }
  Var& operator/=(const Var& r) { return *this = bop(VKJIT_BOP_DIV, r); }

  // Debug (types.rs:115-126): the buffer contents once evaluated, the node otherwise
  std::string repr() const {
    const vkjit_var id = id_;
    const std::string s = detail::text_of([id](char* b, size_t c, size_t* l) { return vkjit_var_repr(detail::ir(), id, b, c, l); });
    return is_buffer() ? "Var(\"" + s + "\")" : "Var(" + s + ")";
  }

  // ---- extensions (no reference counterpart) ---------------------------------------------------------
  Var bop(int32_t kind, const Var& r) const {
    vkjit_var out = 0;
    detail::check(vkjit_bop(detail::ir(), kind, id_, r.id(), &out));
    return from_id(out);
  }
  Var uop(int32_t kind) const {
    vkjit_var out = 0;
    detail::check(vkjit_uop(detail::ir(), kind, id_, &out));
    return from_id(out);
  }
  Var cast(VarType ty) const {  // Ir::cast returns the operand itself when the types match (internal.rs:283-290)
    vkjit_var out = 0;
    detail::check(vkjit_cast(detail::ir(), id_, ty, &out));
    if (out == id_) detail::check(vkjit_inc_ref(detail::ir(), out));
    return from_id(out);
  }
  Var reduce(int32_t red) const {
    vkjit_var out = 0;
    detail::check(vkjit_reduce(detail::ir(), red, id_, &out));
    return from_id(out);
  }
  Var sum() const { return reduce(VKJIT_RED_SUM); }
  Var min() const { return reduce(VKJIT_RED_MIN); }
  Var max() const { return reduce(VKJIT_RED_MAX); }
  Var prefix_sum(bool exclusive = true) const {
    vkjit_var out = 0;
    detail::check(vkjit_prefix_sum(detail::ir(), id_, exclusive ? 1 : 0, &out));
    return from_id(out);
  }
  void scatter_add(const Var& to, const Var& idx) {
    vkjit_var out = 0;
    detail::check(vkjit_scatter_add(detail::ir(), id_, to.id(), idx.id(), 0, 0, &out));
    *this = from_id(out);
  }

 private:
  struct Adopt {};
  Var(vkjit_var id, Adopt) : id_(id) {}
  void scatter_impl(const Var& to, const Var& idx, const Var* cond) {
    vkjit_var out = 0;
    detail::check(vkjit_scatter(detail::ir(), id_, to.id(), idx.id(), cond ? 1 : 0, cond ? cond->id() : 0, &out));
    *this = from_id(out);
  }
  vkjit_var id_ = 0;
  bool live_ = true;
};

inline Var operator+(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_ADD, b); }
inline Var operator-(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_SUB, b); }
inline Var operator*(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_MUL, b); }
inline Var operator/(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_DIV, b); }
inline Var operator&(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_AND, b); }
inline Var operator|(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_OR, b); }
inline Var operator^(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_XOR, b); }
inline Var operator<<(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_SHL, b); }
inline Var operator>>(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_SHR, b); }
inline std::ostream& operator<<(std::ostream& os, const Var& v) { return os << v.repr(); }

// ---- free functions (functions.rs:5-56) ---------------------------------------------------------------------------------
inline Var zeros(VarType ty) {
  vkjit_var out = 0;
  detail::check(vkjit_zeros(detail::ir(), ty, &out));
  return Var::from_id(out);
}
inline Var arange(VarType ty, size_t num) {
  vkjit_var out = 0;
  detail::check(vkjit_arange(detail::ir(), ty, num, &out));
  return Var::from_id(out);
}
inline Var linspace(const Var& start, const Var& stop, size_t num) {  // uses start.ty() (functions.rs:13-22)
  vkjit_var out = 0;
  detail::check(vkjit_linspace(detail::ir(), start.ty(), start.id(), stop.id(), num, &out));
  return Var::from_id(out);
}
inline Var select(const Var& condition, const Var& x, const Var& y) { return condition.then_else(x, y); }
inline Var gather(const Var& from, const Var& idx) {
  vkjit_var out = 0;
  detail::check(vkjit_gather(detail::ir(), from.id(), idx.id(), 0, 0, &out));
  return Var::from_id(out);
}
inline Var gather_with(const Var& from, const Var& idx, const Var& condition) {
  vkjit_var out = 0;
  detail::check(vkjit_gather(detail::ir(), from.id(), idx.id(), 1, condition.id(), &out));
  return Var::from_id(out);
}
inline Var Var::get(const Var& idx) const { return gather(*this, idx); }
inline std::string repr_ir() {
  return detail::text_of([](char* b, size_t c, size_t* l) { return vkjit_ir_repr(detail::ir(), b, c, l); });
}

// eval!(a, b, ...) / schedule!(a, b, ...) (functions.rs:58-82)
inline void eval_internal(const std::vector<vkjit_var>& schedule) { detail::check(vkjit_eval(detail::ir(), schedule.data(), schedule.size())); }
inline void schedule_internal(const std::vector<vkjit_var>& schedule) {
  detail::check(vkjit_schedule(detail::ir(), schedule.data(), schedule.size()));
}
template <class... Vars> void eval(const Vars&... vars) { eval_internal(std::vector<vkjit_var>{vars.id()...}); }
template <class... Vars> void schedule(const Vars&... vars) { schedule_internal(std::vector<vkjit_var>{vars.id()...}); }

// ---- extensions ---------------------------------------------------------------------------------------------------------------
inline Var sqrt(const Var& x) { return x.uop(VKJIT_UOP_SQRT); }
inline Var exp(const Var& x) { return x.uop(VKJIT_UOP_EXP); }
inline Var log(const Var& x) { return x.uop(VKJIT_UOP_LOG); }
inline Var sin(const Var& x) { return x.uop(VKJIT_UOP_SIN); }
inline Var cos(const Var& x) { return x.uop(VKJIT_UOP_COS); }
inline Var minimum(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_MIN, b); }
inline Var maximum(const Var& a, const Var& b) { return a.bop(VKJIT_BOP_MAX, b); }
// stream compaction: values[mask] in order, and their count
inline std::pair<Var, size_t> compress(const Var& values, const Var& mask) {
  vkjit_var out = 0;
  size_t count = 0;
  detail::check(vkjit_compress_values(detail::ir(), values.id(), mask.id(), &out, &count));
  return {Var::from_id(out), count};
}
inline void sync() { detail::check(vkjit_sync()); }

}  // namespace vkjit
