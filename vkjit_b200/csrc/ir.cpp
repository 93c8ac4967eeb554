// ir.cpp — trace construction, type promotion and ref-counting.
// Reference behaviour being matched: libs/vkjit-core/src/internal.rs:146-481, vartype.rs:24-83.
//
// Intentional deviations from the reference (documented in DESIGN.md):
//  * implicit casts created by bop!, linspace temporaries and ones(Struct) members do not
//    keep their initial reference (the reference never releases them: internal.rs:159-160,
//    :238-246, :267-275 — a leak that pins device arrays forever);
//  * releasing a var releases every dependency EDGE (the reference's MutSeVisitor skips a
//    dependency that occurs twice, iterators.rs:69-87 — another leak);
//  * slots of dead vars are recycled.
#include "ir.h"

#include <algorithm>
#include <charconv>
#include <cmath>

namespace vkjit {

Ir::Ir() { vars.reserve(1024); }

Ir::~Ir() {
  for (Var& v : vars) {
    if (v.array) { release_array(v.array); v.array = nullptr; }
  }
}

Var& Ir::var(VarId id) {
  if (id >= vars.size() || vars[id].op == OP_FREE) fail(VKJIT_ERR_INVALID, "invalid VarId " + std::to_string(id));
  return vars[id];
}
const Var& Ir::var(VarId id) const {
  if (id >= vars.size() || vars[id].op == OP_FREE) fail(VKJIT_ERR_INVALID, "invalid VarId " + std::to_string(id));
  return vars[id];
}

uint32_t Ir::next_stamp() {
  if (++stamp_counter == 0) {  // wrapped: reset all scratch stamps
    for (Var& v : vars) v.stamp = 0;
    stamp_counter = 1;
  }
  return stamp_counter;
}

// ---- types -------------------------------------------------------------------
void Ir::check_type(TypeId t) const {
  if (t >= VKJIT_TY_VOID && t <= VKJIT_TY_F32) return;
  if (t >= VKJIT_TY_STRUCT_BASE && t - VKJIT_TY_STRUCT_BASE < struct_types.size()) return;
  fail(VKJIT_ERR_INVALID, "invalid type code " + std::to_string(t));
}

TypeId Ir::struct_type(const TypeId* elems, size_t n) {
  std::vector<TypeId> e(elems, elems + n);
  for (TypeId t : e) check_type(t);
  for (size_t i = 0; i < struct_types.size(); ++i)
    if (struct_types[i] == e) return VKJIT_TY_STRUCT_BASE + (TypeId)i;
  struct_types.push_back(std::move(e));
  return VKJIT_TY_STRUCT_BASE + (TypeId)(struct_types.size() - 1);
}

const std::vector<TypeId>& Ir::struct_elems(TypeId t) const {
  if (!ty_is_struct(t) || t - VKJIT_TY_STRUCT_BASE >= struct_types.size()) fail(VKJIT_ERR_TYPE, "not a struct type");
  return struct_types[t - VKJIT_TY_STRUCT_BASE];
}

// vartype.rs:24-33: Struct < Void < Bool < U32 < I32 < F32 (derive(Ord)); max returns rhs on ties.
TypeId Ir::ty_max(TypeId a, TypeId b) {
  const uint32_t ra = ty_is_struct(a) ? 0 : a, rb = ty_is_struct(b) ? 0 : b;
  return ra > rb ? a : b;
}

// ---- slots ---------------------------------------------------------------------
VarId Ir::alloc_slot() {
  if (!free_list.empty()) {
    VarId id = free_list.back();
    free_list.pop_back();
    return id;
  }
  vars.emplace_back();
  return (VarId)(vars.size() - 1);
}

// internal.rs:186-209: ref_count = 1, every dependency +1
VarId Ir::new_var(Op op, TypeId ty, const VarId* deps, size_t ndeps, uint16_t kind, uint32_t aux) {
  uint8_t sharded = 0;
  for (size_t i = 0; i < ndeps; ++i) sharded |= var(deps[i]).sharded;
  VarId id = alloc_slot();
  Var& v = vars[id];
  v = Var();
  v.op = op; v.ty = ty; v.kind = kind; v.aux = aux; v.ref_count = 1; v.sharded = sharded;
  v.ndeps = (uint32_t)ndeps;
  if (ndeps <= 3) { for (size_t i = 0; i < ndeps; ++i) v.dep_inline[i] = deps[i]; }
  else { v.dep_ext.reset(new VarId[ndeps]); for (size_t i = 0; i < ndeps; ++i) v.dep_ext[i] = deps[i]; }
  for (size_t i = 0; i < ndeps; ++i) vars[deps[i]].ref_count += 1;
  return id;
}

void Ir::inc_ref(VarId id) { var(id).ref_count += 1; }  // internal.rs:466-469

// internal.rs:450-465: at zero the array is dropped and the release cascades.  Only vars that actually die go
// on the work stack (a consumed trace releases every node: ~5 ns per node instead of ~8).
void Ir::dec_ref(VarId id) {
  Var& r = var(id);
  if (r.op == OP_FREE || r.ref_count == 0) fail(VKJIT_ERR_INVALID, "ref_count underflow on var " + std::to_string(id));
  if (--r.ref_count != 0) return;
  dec_stack_.clear();
  dec_stack_.push_back(id);
  Var* const vs = vars.data();
  auto release_edge = [&](VarId d) {
    Var& dv = vs[d];
    if (dv.op == OP_FREE || dv.ref_count == 0) fail(VKJIT_ERR_INVALID, "ref_count underflow on var " + std::to_string(d));
    if (--dv.ref_count == 0) dec_stack_.push_back(d);
  };
  while (!dec_stack_.empty()) {
    const VarId cur = dec_stack_.back();
    dec_stack_.pop_back();
    Var& v = vs[cur];
    if (v.array) { release_array(v.array); v.array = nullptr; --n_arrays; }
    const VarId* d = v.deps();
    for (uint32_t i = 0; i < v.ndeps; ++i) release_edge(d[i]);
    if (v.has_se) release_edge(v.side_effect);
    v.ndeps = 0; v.has_se = false;
    if (v.dep_ext) v.dep_ext.reset();
    v.op = OP_FREE;  // ref_count stays readable as 0 (test.rs:205)
    free_list.push_back(cur);
  }
}

// ---- constructors ----------------------------------------------------------------
VarId Ir::constant(TypeId ty, uint32_t bits) { return new_var(OP_CONST, ty, nullptr, 0, 0, bits); }

VarId Ir::binding(TypeId ty, Array* arr, bool sharded) {
  VarId id = new_var(OP_BINDING, ty, nullptr, 0);
  vars[id].array = arr;
  vars[id].sharded = sharded;
  ++n_arrays;
  return id;
}

// internal.rs:235-237
VarId Ir::arange(TypeId ty, uint64_t n, uint64_t base, bool sharded) {
  if (!ty_is_num(ty)) fail(VKJIT_ERR_UNSUPPORTED, "arange of this type (reference: unimplemented!(), internal.rs:1090)");
  if (n > 0xFFFFFFFFull || base + n > 0x100000000ull) fail(VKJIT_ERR_SIZE, "arange exceeds the 32-bit lane index");
  VarId id = new_var(OP_ARANGE, ty, nullptr, 0);
  vars[id].num = n; vars[id].base = base; vars[id].sharded = sharded;
  return id;
}

// internal.rs:283-290
VarId Ir::cast(VarId src, TypeId ty) {
  check_type(ty);
  const TypeId st = var(src).ty;
  if (st == ty) return src;
  if (!ty_is_scalar(st) || !ty_is_scalar(ty)) fail(VKJIT_ERR_UNSUPPORTED, "cast of a non-scalar type");
  return new_var(OP_CAST, ty, &src, 1);
}

// internal.rs:146-166
VarId Ir::bop(int kind, VarId lhs, VarId rhs) {
  const TypeId lt = var(lhs).ty, rt = var(rhs).ty;
  const bool ref_kind = kind >= VKJIT_BOP_ADD && kind <= VKJIT_BOP_NEQ;
  const bool ext_kind = kind >= VKJIT_BOP_AND && kind <= VKJIT_BOP_MAX;
  if (!ref_kind && !ext_kind) fail(VKJIT_ERR_INVALID, "unknown binary op");
  if (!ty_is_scalar(lt) || !ty_is_scalar(rt)) fail(VKJIT_ERR_UNSUPPORTED, "binary op on a non-scalar type");
  const TypeId opty = ty_max(lt, rt);
  const bool cmp = kind >= VKJIT_BOP_LT && kind <= VKJIT_BOP_NEQ;
  const bool logic = kind == VKJIT_BOP_AND || kind == VKJIT_BOP_OR || kind == VKJIT_BOP_XOR;
  const bool shift = kind == VKJIT_BOP_SHL || kind == VKJIT_BOP_SHR;
  if (opty == VKJIT_TY_BOOL && !logic)
    fail(VKJIT_ERR_UNSUPPORTED, "arithmetic/compare on Bool operands (reference: unimplemented!(), internal.rs:918)");
  if ((logic || shift) && opty == VKJIT_TY_F32) fail(VKJIT_ERR_TYPE, "bit op on F32");
  if (shift && opty == VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "shift on Bool");
  const TypeId ty = cmp ? (TypeId)VKJIT_TY_BOOL : opty;  // Bop::eval_ty, internal.rs:46-52
  const VarId l = cast(lhs, opty), r = cast(rhs, opty);
  const VarId d[2] = {l, r};
  const VarId out = new_var(OP_BOP, ty, d, 2, (uint16_t)kind);
  if (l != lhs) dec_ref(l);  // deviation: the implicit cast is owned by the op alone
  if (r != rhs) dec_ref(r);
  return out;
}

VarId Ir::uop(int kind, VarId src) {
  const TypeId ty = var(src).ty;
  if (!ty_is_scalar(ty)) fail(VKJIT_ERR_UNSUPPORTED, "unary op on a non-scalar type");
  switch (kind) {
    case VKJIT_UOP_NEG: case VKJIT_UOP_ABS:
      if (ty == VKJIT_TY_BOOL) fail(VKJIT_ERR_TYPE, "neg/abs on Bool");
      break;
    case VKJIT_UOP_NOT:
      if (ty == VKJIT_TY_F32) fail(VKJIT_ERR_TYPE, "not on F32");
      break;
    case VKJIT_UOP_SQRT: case VKJIT_UOP_EXP: case VKJIT_UOP_LOG: case VKJIT_UOP_SIN: case VKJIT_UOP_COS:
      if (ty != VKJIT_TY_F32) fail(VKJIT_ERR_TYPE, "transcendental on a non-F32 type");
      break;
    default: fail(VKJIT_ERR_INVALID, "unknown unary op");
  }
  return new_var(OP_UOP, ty, &src, 1, (uint16_t)kind);
}

VarId Ir::bitcast(VarId src, TypeId ty) {
  const TypeId st = var(src).ty;
  if (!ty_is_num(st) || !ty_is_num(ty)) fail(VKJIT_ERR_TYPE, "bitcast only between U32/I32/F32");
  if (st == ty) return src;
  return new_var(OP_BITCAST, ty, &src, 1);
}

// internal.rs:229-234
VarId Ir::select(VarId c, VarId l, VarId r) {
  var(c);
  if (var(l).ty != var(r).ty) fail(VKJIT_ERR_TYPE, "select: lhs and rhs types differ (internal.rs:232)");
  const VarId d[3] = {c, l, r};
  return new_var(OP_SELECT, var(l).ty, d, 3);
}

// Temporaries of a composite constructor: owned by the expression that consumes them, released on EVERY exit — also
// when a later step throws (a Bool / struct `ty`, n >= 2^32, a Void member): nothing stays pinned by a failed call.
namespace {
struct Temps {
  Ir& ir;
  std::vector<VarId> ids;
  explicit Temps(Ir& i) : ir(i) {}
  VarId keep(VarId id) { ids.push_back(id); return id; }
  ~Temps() { for (size_t k = ids.size(); k-- > 0;) ir.dec_ref(ids[k]); }
};
}  // namespace

// internal.rs:238-246: ((arange(ty, n) / u32(n)) * (stop - start)) + start
VarId Ir::linspace(TypeId ty, VarId start, VarId stop, uint64_t n) {
  var(start); var(stop);
  Temps t(*this);  // deviation: the reference leaks these five vars (internal.rs:238-246)
  const VarId len = t.keep(bop(VKJIT_BOP_SUB, stop, start));
  const VarId idx = t.keep(arange(ty, n, 0, false));
  const VarId num = t.keep(constant(VKJIT_TY_U32, (uint32_t)n));
  const VarId a = t.keep(bop(VKJIT_BOP_DIV, idx, num));
  const VarId b = t.keep(bop(VKJIT_BOP_MUL, a, len));
  return bop(VKJIT_BOP_ADD, b, start);
}

// internal.rs:291-300
VarId Ir::struct_init(const VarId* elems, size_t n) {
  std::vector<TypeId> tys(n);
  for (size_t i = 0; i < n; ++i) tys[i] = var(elems[i]).ty;
  const TypeId st = struct_type(tys.data(), n);
  return new_var(OP_STRUCTINIT, st, elems, n);
}

static uint32_t one_bits(TypeId t) { return t == VKJIT_TY_F32 ? 0x3F800000u : 1u; }

// internal.rs:247-266
VarId Ir::zeros(TypeId ty) {
  check_type(ty);
  if (ty_is_struct(ty)) {
    const std::vector<TypeId> elems = struct_elems(ty);
    Temps t(*this);
    for (TypeId e : elems) t.keep(zeros(e));
    return struct_init(t.ids.data(), t.ids.size());
  }
  if (!ty_is_scalar(ty)) fail(VKJIT_ERR_UNSUPPORTED, "zeros of Void");
  return constant(ty, 0u);
}

// internal.rs:267-282
VarId Ir::ones(TypeId ty) {
  check_type(ty);
  if (ty_is_struct(ty)) {
    const std::vector<TypeId> elems = struct_elems(ty);
    Temps t(*this);  // deviation: see file header
    for (TypeId e : elems) t.keep(ones(e));
    return struct_init(t.ids.data(), t.ids.size());
  }
  if (!ty_is_scalar(ty)) fail(VKJIT_ERR_UNSUPPORTED, "ones of Void");
  return constant(ty, one_bits(ty));
}

// internal.rs:349-356
VarId Ir::getattr(VarId src, size_t idx) {
  const TypeId t = var(src).ty;
  if (!ty_is_struct(t)) fail(VKJIT_ERR_UNSUPPORTED, "getattr on a non-struct (internal.rs:353)");
  const auto& e = struct_elems(t);
  if (idx >= e.size()) fail(VKJIT_ERR_INVALID, "getattr index out of range");
  return new_var(OP_GETATTR, e[idx], &src, 1, 0, (uint32_t)idx);
}

// internal.rs:357-367 — deps = [src, dst]
VarId Ir::setattr(VarId dst, VarId src, size_t idx) {
  const TypeId t = var(dst).ty;
  if (!ty_is_struct(t)) fail(VKJIT_ERR_UNSUPPORTED, "setattr on a non-struct");
  const auto& e = struct_elems(t);
  if (idx >= e.size()) fail(VKJIT_ERR_INVALID, "setattr index out of range");
  if (e[idx] != var(src).ty) fail(VKJIT_ERR_TYPE, "setattr: member type mismatch");
  const VarId d[2] = {src, dst};
  return new_var(OP_SETATTR, t, d, 2, 0, (uint32_t)idx);
}

// internal.rs:368-378 — deps = [src, idx(, active)]
VarId Ir::gather(VarId src, VarId idx, bool has_active, VarId active) {
  const TypeId t = var(src).ty;
  var(idx);
  VarId d[3] = {src, idx, 0};
  size_t n = 2;
  if (has_active) { var(active); d[2] = active; n = 3; }
  const VarId out = new_var(OP_GATHER, t, d, n);
  // the gathered array is addressed by index, not by lane: a replicated source does not make
  // the result lane-sharded, only idx/active do
  vars[out].sharded = vars[idx].sharded | (has_active ? vars[active].sharded : 0);
  return out;
}

// internal.rs:379-400 — deps = [src, idx(, active)], side_effects = [dst]
VarId Ir::scatter(Op op, VarId src, VarId dst, VarId idx, bool has_active, VarId active) {
  const TypeId t = var(src).ty;
  var(idx); var(dst);
  VarId d[3] = {src, idx, 0};
  size_t n = 2;
  if (has_active) { var(active); d[2] = active; n = 3; }
  const VarId out = new_var(op, t, d, n);
  vars[out].has_se = true;
  vars[out].side_effect = dst;
  vars[dst].ref_count += 1;
  return out;
}

// ---- schedule / eval bookkeeping -----------------------------------------------------
// internal.rs:476-481
void Ir::do_schedule(const VarId* ids, size_t n) {
  for (size_t i = 0; i < n; ++i) var(ids[i]);
  for (size_t i = 0; i < n; ++i) {
    if (std::find(schedule.begin(), schedule.end(), ids[i]) != schedule.end()) continue;  // duplicate roots collapse
    vars[ids[i]].ref_count += 1;
    schedule.push_back(ids[i]);
  }
}

// internal.rs:470-475
void Ir::clear_schedule() {
  std::vector<VarId> s;
  s.swap(schedule);
  for (VarId id : s) dec_ref(id);
}

// internal.rs:492-521
void Ir::commit_roots(const std::vector<VarId>& roots, const std::vector<Array*>& outs, const std::vector<VarId>& order) {
  // 1) every root gives up its dependency edges and is rewritten into a Binding owning its output, 2) the
  // vars those edges kept alive are released.  Same end state as the reference's cascade (dec_ref_count per
  // dependency, internal.rs:497-509).  `order` is the post-order of the trace that just ran: every consumer
  // comes after its operands, so ONE reverse sweep releases the whole trace without a work stack
  // (a consumed 364-node trace: 1.5 -> ~1 us).
  Var* const vs = vars.data();
  auto release_edge = [&](VarId d) {
    Var& dv = vs[d];
    if (dv.op == OP_FREE || dv.ref_count == 0) fail(VKJIT_ERR_INVALID, "ref_count underflow on var " + std::to_string(d));
    --dv.ref_count;
  };
  for (size_t i = 0; i < roots.size(); ++i) {
    Var& v = vs[roots[i]];
    const VarId* d = v.deps();
    for (uint32_t k = 0; k < v.ndeps; ++k) release_edge(d[k]);
    if (v.has_se) release_edge(v.side_effect);
    if (v.array) { release_array(v.array); v.array = nullptr; --n_arrays; }  // a root that already was a Binding is copied (internal.rs:1192-1205)
    v.op = OP_BINDING;
    v.kind = 0; v.aux = 0; v.num = 0; v.base = 0;
    v.ndeps = 0; v.has_se = false;
    if (v.dep_ext) v.dep_ext.reset();
    v.array = outs[i];
    ++n_arrays;
    // ty, ref_count and the sharded flag are kept
  }
  for (size_t i = order.size(); i-- > 0;) {
    const VarId id = order[i];
    Var& v = vs[id];
    if (v.ref_count != 0 || v.op == OP_FREE) continue;
    if (v.array) { release_array(v.array); v.array = nullptr; --n_arrays; }
    const VarId* d = v.deps();
    for (uint32_t k = 0; k < v.ndeps; ++k) release_edge(d[k]);
    if (v.has_se) release_edge(v.side_effect);
    v.ndeps = 0; v.has_se = false;
    if (v.dep_ext) v.dep_ext.reset();
    v.op = OP_FREE;  // ref_count stays readable as 0 (test.rs:205)
    free_list.push_back(id);
  }
}

// ---- Debug output ---------------------------------------------------------------------
std::string format_f32(float f) {
  if (f != f) return "NaN";
  if (std::isinf(f)) return f < 0 ? "-inf" : "inf";
  if (f == 0.f) return std::signbit(f) ? "-0.0" : "0.0";
  char buf[48];
  auto r = std::to_chars(buf, buf + sizeof buf, std::fabs(f), std::chars_format::scientific);
  std::string s(buf, r.ptr);
  const size_t e = s.find('e');
  const int exp = atoi(s.c_str() + e + 1);
  std::string digits;
  for (size_t i = 0; i < e; ++i) if (s[i] != '.') digits.push_back(s[i]);
  std::string out;
  if (exp >= -5 && exp < 16) {
    if (exp < 0) out = "0." + std::string((size_t)(-exp - 1), '0') + digits;
    else if ((int)digits.size() <= exp + 1) out = digits + std::string((size_t)(exp + 1) - digits.size(), '0') + ".0";
    else out = digits.substr(0, (size_t)exp + 1) + "." + digits.substr((size_t)exp + 1);
  } else {
    out = digits.substr(0, 1) + (digits.size() > 1 ? "." + digits.substr(1) : "") + "e" + std::to_string(exp);
  }
  return f < 0 ? "-" + out : out;
}

std::string Ir::type_name(TypeId t) const {
  switch (t) {
    case VKJIT_TY_VOID: return "Void";
    case VKJIT_TY_BOOL: return "Bool";
    case VKJIT_TY_U32: return "U32";
    case VKJIT_TY_I32: return "I32";
    case VKJIT_TY_F32: return "F32";
    default: break;
  }
  std::string s = "Struct([";
  const auto& e = struct_elems(t);
  for (size_t i = 0; i < e.size(); ++i) s += (i ? ", " : "") + type_name(e[i]);
  return s + "])";
}

namespace {
const char* kBop[] = {"Add", "Sub", "Mul", "Div", "Lt", "Gt", "Eq", "Leq", "Geq", "Neq"};
const char* kBopExt[] = {"And", "Or", "Xor", "Shl", "Shr", "Min", "Max"};
const char* kUop[] = {"Neg", "Abs", "Not", "Sqrt", "Exp", "Log", "Sin", "Cos"};

// Writes either the one-line `{:?}` or the indented `{:#?}` form.
struct DebugWriter {
  bool pretty;
  int depth = 0;
  std::string out;
  void nl() { if (pretty) { out += "\n"; out.append((size_t)depth * 4, ' '); } }
  void open(const std::string& head, char br) { out += head; out += br; ++depth; }
  void close(char br, bool had_items) {
    --depth;
    if (had_items && pretty) { out += ",\n"; out.append((size_t)depth * 4, ' '); }
    out += br;
  }
  void tuple1(const std::string& name, const std::string& atom) {
    open(name, '('); nl(); out += atom; close(')', true);
  }
};
}  // namespace

static void write_type(const Ir& ir, TypeId t, DebugWriter& w) {
  if (!ty_is_struct(t)) { w.out += ir.type_name(t); return; }
  const auto& e = ir.struct_elems(t);
  w.open("Struct", '('); w.nl();
  w.open("", '[');
  for (size_t i = 0; i < e.size(); ++i) {
    if (i) w.out += w.pretty ? "," : ", ";
    w.nl(); write_type(ir, e[i], w);
  }
  w.close(']', !e.empty());
  w.close(')', true);
}

static void write_ids(const std::vector<VarId>& ids, DebugWriter& w) {
  w.open("", '[');
  for (size_t i = 0; i < ids.size(); ++i) {
    if (i) w.out += w.pretty ? "," : ", ";
    w.nl(); w.out += std::to_string(ids[i]);
  }
  w.close(']', !ids.empty());
}

static void write_var(const Ir& ir, const Var& v, DebugWriter& w) {
  auto field = [&](const char* name, bool first) {
    if (!first) w.out += w.pretty ? "," : ", ";
    if (w.pretty) w.nl();
    w.out += name; w.out += ": ";
  };
  w.out += "Var";
  w.out += w.pretty ? " {" : " { ";
  ++w.depth;
  field("op", true);
  switch (v.op) {
    case OP_FREE: w.out += "Free"; break;
    case OP_BINDING: w.out += "Binding"; break;
    case OP_BOP: w.tuple1("Bop", v.kind < 16 ? kBop[v.kind] : kBopExt[v.kind - 16]); break;
    case OP_ARANGE: w.tuple1("Arange", std::to_string(v.num)); break;
    case OP_CONST: {
      w.open("Const", '('); w.nl();
      switch (v.ty) {
        case VKJIT_TY_BOOL: w.tuple1("Bool", v.aux ? "true" : "false"); break;
        case VKJIT_TY_U32: w.tuple1("UInt32", std::to_string(v.aux)); break;
        case VKJIT_TY_I32: w.tuple1("Int32", std::to_string((int32_t)v.aux)); break;
        default: w.tuple1("Float32", format_f32(bits_f32(v.aux))); break;
      }
      w.close(')', true);
      break;
    }
    case OP_GETATTR: w.tuple1("GetAttr", std::to_string(v.aux)); break;
    case OP_SETATTR: w.tuple1("SetAttr", std::to_string(v.aux)); break;
    case OP_STRUCTINIT: w.out += "StructInit"; break;
    case OP_GATHER: w.out += "Gather"; break;
    case OP_SCATTER: w.out += "Scatter"; break;
    case OP_SELECT: w.out += "Select"; break;
    case OP_CAST: w.out += "Cast"; break;
    case OP_UOP: w.tuple1("Uop", kUop[v.kind]); break;
    case OP_BITCAST: w.out += "Bitcast"; break;
    case OP_SCATTER_ADD: w.out += "ScatterAdd"; break;
  }
  field("deps", false);
  write_ids(std::vector<VarId>(v.deps(), v.deps() + v.ndeps), w);
  field("side_effects", false);
  write_ids(v.has_se ? std::vector<VarId>{v.side_effect} : std::vector<VarId>{}, w);
  field("ty", false);
  write_type(ir, v.ty, w);
  field("ref_count", false);
  w.out += std::to_string(v.ref_count);
  --w.depth;
  if (w.pretty) { w.out += ",\n"; w.out.append((size_t)w.depth * 4, ' '); w.out += "}"; }
  else w.out += " }";
}

std::string Ir::var_debug(VarId id) const {
  DebugWriter w{false};
  write_var(*this, var(id), w);
  return w.out;
}

std::string Ir::repr() const {
  if (vars.empty()) return "Ir";
  DebugWriter w{true};
  w.out = "Ir {";
  w.depth = 1;
  for (size_t i = 0; i < vars.size(); ++i) {
    w.nl();
    w.out += "[" + std::to_string(i) + "]: ";
    write_var(*this, vars[i], w);
    w.out += ",";
  }
  w.out += "\n}";
  return w.out;
}

}  // namespace vkjit
