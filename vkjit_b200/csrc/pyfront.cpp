// pyfront.cpp — native half of the `vkjit` Python module (vkjit_b200/_native.*.so).
//
// The reference's Python front-end is a compiled pyo3 module (libs/vkjit-python/src/{lib,types,functions}.rs):
// `Var` is a #[pyclass] over a VarId, its operators call straight into `Ir`.  This file is the same thing over
// the C ABI (include/vkjit_b200.h): the `Var` base type with the operator slots, the argument coercion of
// types.rs:47-82 for scalars, and the handful of module functions a trace is built from.  Everything that is not
// on the trace-building path (readback, NumPy / DLPack interop, scatter, reductions) stays in
// vkjit_b200/vkjit.py, which subclasses this type.  Measured on the 364-node Monte-Carlo trace: building the
// trace costs ~0.2 us per node here instead of ~3 us through ctypes.
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/vkjit_b200.h"

namespace {

vkjit_ir* g_ir = nullptr;            // handle of the global Ir (`lazy_static IR`, vkjit-python/src/lib.rs:14-16)
PyObject* g_ir_owner = nullptr;      // the Python object owning that handle: kept alive while Vars may still release
PyObject* g_slow_coerce = nullptr;   // Python callable: sequences / NumPy values -> Var (uploads need NumPy)
PyObject* g_lazy_bind = nullptr;     // Python callable: creates the global Ir and calls bind()
PyTypeObject* g_var_cls = nullptr;   // class of the objects the operators return (the Python subclass)
PyObject* g_errs[4] = {nullptr, nullptr, nullptr, nullptr};  // VkjitError, VkjitTypeError, VkjitSizeError, VkjitNoDeviceError

struct VarObject {
  PyObject_HEAD
  uint32_t id;
  int live;      // owns one reference count of `id` (Clone / Drop, types.rs:87-98)
  uint32_t gen;  // binding generation it was created under: a handle of a closed Ir never touches its successor
};
uint32_t g_gen = 0;  // bumped by every bind()

extern PyTypeObject VarBase_Type;

PyObject* raise_status(vkjit_status st) {
  const char* msg = vkjit_last_error();
  if (!msg) msg = "";
  PyObject* cls = st == VKJIT_ERR_TYPE ? g_errs[1] : st == VKJIT_ERR_SIZE ? g_errs[2] : st == VKJIT_ERR_NO_DEVICE ? g_errs[3] : g_errs[0];
  if (!cls) { PyErr_Format(PyExc_RuntimeError, "[status %d] %s", (int)st, msg); return nullptr; }
  PyObject* text = PyUnicode_DecodeUTF8(msg, (Py_ssize_t)strlen(msg), "replace");
  if (!text) return nullptr;
  PyObject* exc = PyObject_CallFunction(cls, "iN", (int)st, text);
  if (exc) { PyErr_SetObject(cls, exc); Py_DECREF(exc); }
  return nullptr;
}

bool ensure_ir() {
  if (g_ir) return true;
  if (!g_lazy_bind) { PyErr_SetString(PyExc_RuntimeError, "vkjit front-end is not configured"); return false; }
  PyObject* r = PyObject_CallNoArgs(g_lazy_bind);
  if (!r) return false;
  Py_DECREF(r);
  if (!g_ir) { PyErr_SetString(PyExc_RuntimeError, "vkjit front-end: no Ir bound"); return false; }
  return true;
}

PyObject* invalid_argument() {
  PyErr_SetString(PyExc_TypeError, "Not a valid argument!");  // types.rs:81
  return nullptr;
}

PyObject* wrap(PyTypeObject* cls, vkjit_var id) {
  VarObject* o = (VarObject*)cls->tp_alloc(cls, 0);
  if (!o) { vkjit_dec_ref(g_ir, id); return nullptr; }
  o->id = id; o->live = 1; o->gen = g_gen;
  return (PyObject*)o;
}
inline PyObject* wrap(vkjit_var id) { return wrap(g_var_cls ? g_var_cls : &VarBase_Type, id); }

// One operand of an op: the id, whether this call created it (a const for a Python scalar: the op's own reference
// keeps it alive, ours is dropped afterwards) and an owner object to release (slow path).
struct Operand {
  vkjit_var id = 0;
  bool temp = false;
  PyObject* keep = nullptr;
};

void release(Operand& o) {
  if (o.temp) vkjit_dec_ref(g_ir, o.id);
  Py_XDECREF(o.keep);
  o.temp = false; o.keep = nullptr;
}

// TryFrom<&PyAny> for Var (types.rs:47-82): Var, u32, i32, f32, bool, [u32], [i32], [f32].  pyo3 tries u32 first, and a
// Python bool is an int, so True becomes UInt32(1).
bool coerce(PyObject* v, Operand& o) {
  vkjit_status st;
  if (PyObject_TypeCheck(v, &VarBase_Type)) {
    VarObject* x = (VarObject*)v;
    if (!x->live || x->gen != g_gen) { PyErr_SetString(PyExc_TypeError, "this Var no longer owns a variable"); return false; }
    o.id = x->id;
    return true;
  }
  if (PyBool_Check(v)) {
    st = vkjit_const_u32(g_ir, v == Py_True ? 1u : 0u, &o.id);
  } else if (PyLong_Check(v)) {
    int overflow = 0;
    const long long x = PyLong_AsLongLongAndOverflow(v, &overflow);
    if (x == -1 && !overflow && PyErr_Occurred()) return false;
    if (overflow || x > 0xFFFFFFFFLL || x < -2147483648LL) { invalid_argument(); return false; }
    st = x >= 0 ? vkjit_const_u32(g_ir, (uint32_t)x, &o.id) : vkjit_const_i32(g_ir, (int32_t)x, &o.id);
  } else if (PyFloat_Check(v)) {
    st = vkjit_const_f32(g_ir, (float)PyFloat_AS_DOUBLE(v), &o.id);
  } else {
    if (!g_slow_coerce) { invalid_argument(); return false; }
    PyObject* r = PyObject_CallOneArg(g_slow_coerce, v);
    if (!r) return false;
    if (!PyObject_TypeCheck(r, &VarBase_Type) || !((VarObject*)r)->live || ((VarObject*)r)->gen != g_gen) {
      Py_DECREF(r);
      invalid_argument();
      return false;
    }
    o.id = ((VarObject*)r)->id;
    o.keep = r;
    return true;
  }
  if (st != VKJIT_OK) { raise_status(st); return false; }
  o.temp = true;
  return true;
}

// A new owned Var object for any accepted value (a Var argument is cloned).
PyObject* coerce_to_object(PyObject* v) {
  Operand o;
  if (!coerce(v, o)) return nullptr;
  if (o.keep) return o.keep;  // the slow path already made one
  if (!o.temp) {
    const vkjit_status st = vkjit_inc_ref(g_ir, o.id);
    if (st != VKJIT_OK) return raise_status(st);
  }
  return wrap(o.id);
}

PyObject* do_bop(int kind, PyObject* a, PyObject* b) {
  if (!ensure_ir()) return nullptr;
  Operand x, y;
  if (!coerce(a, x)) return nullptr;
  if (!coerce(b, y)) { release(x); return nullptr; }
  vkjit_var out = 0;
  const vkjit_status st = vkjit_bop(g_ir, kind, x.id, y.id, &out);
  release(x); release(y);
  if (st != VKJIT_OK) return raise_status(st);
  return wrap(out);
}

PyObject* do_uop(int kind, PyObject* a) {
  if (!ensure_ir()) return nullptr;
  Operand x;
  if (!coerce(a, x)) return nullptr;
  vkjit_var out = 0;
  const vkjit_status st = vkjit_uop(g_ir, kind, x.id, &out);
  release(x);
  if (st != VKJIT_OK) return raise_status(st);
  return wrap(out);
}

// ---- number protocol: `impl Add/Sub/Mul/Div for Var` (types.rs:124-139) + the bit operators ------------------------
#define VK_BINARY(name, kind) \
  PyObject* name(PyObject* a, PyObject* b) { return do_bop(kind, a, b); }
VK_BINARY(nb_add, VKJIT_BOP_ADD)
VK_BINARY(nb_sub, VKJIT_BOP_SUB)
VK_BINARY(nb_mul, VKJIT_BOP_MUL)
VK_BINARY(nb_div, VKJIT_BOP_DIV)
VK_BINARY(nb_and, VKJIT_BOP_AND)
VK_BINARY(nb_or, VKJIT_BOP_OR)
VK_BINARY(nb_xor, VKJIT_BOP_XOR)
VK_BINARY(nb_shl, VKJIT_BOP_SHL)
VK_BINARY(nb_shr, VKJIT_BOP_SHR)
#undef VK_BINARY
PyObject* nb_neg(PyObject* a) { return do_uop(VKJIT_UOP_NEG, a); }
PyObject* nb_invert(PyObject* a) { return do_uop(VKJIT_UOP_NOT, a); }
PyObject* nb_abs(PyObject* a) { return do_uop(VKJIT_UOP_ABS, a); }

PyNumberMethods var_as_number = {
    nb_add, nb_sub, nb_mul,
    nullptr,  // nb_remainder
    nullptr,  // nb_divmod
    nullptr,  // nb_power
    nb_neg,
    nullptr,  // nb_positive
    nb_abs,
    nullptr,  // nb_bool
    nb_invert, nb_shl, nb_shr, nb_and, nb_xor, nb_or,
    nullptr,  // nb_int
    nullptr,  // nb_reserved
    nullptr,  // nb_float
    nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,  // in-place ops
    nullptr,  // nb_floor_divide
    nb_div,   // nb_true_divide
    nullptr, nullptr,  // in-place floor/true divide
    nullptr,  // nb_index
    nullptr, nullptr,  // matrix multiply
};

// `<`, `>`, `<=`, `>=` build comparison nodes; `==` / `!=` stay identity (use .eq() / .neq()), so a Var is hashable.
PyObject* var_richcompare(PyObject* a, PyObject* b, int op) {
  switch (op) {
    case Py_LT: return do_bop(VKJIT_BOP_LT, a, b);
    case Py_GT: return do_bop(VKJIT_BOP_GT, a, b);
    case Py_LE: return do_bop(VKJIT_BOP_LEQ, a, b);
    case Py_GE: return do_bop(VKJIT_BOP_GEQ, a, b);
    default: Py_RETURN_NOTIMPLEMENTED;
  }
}

// ---- Var methods ----------------------------------------------------------------------------------------------------
inline VarObject* self_live(PyObject* self) {
  VarObject* v = (VarObject*)self;
  if (!v->live || v->gen != g_gen) { PyErr_SetString(PyExc_TypeError, "this Var no longer owns a variable"); return nullptr; }
  return v;
}

PyObject* var_id(PyObject* self, PyObject*) {
  VarObject* v = (VarObject*)self;
  if (!v->live) Py_RETURN_NONE;
  return PyLong_FromUnsignedLong(v->id);
}

PyObject* var_ty(PyObject* self, PyObject*) {
  VarObject* v = self_live(self);
  if (!v || !ensure_ir()) return nullptr;
  vkjit_type t = 0;
  const vkjit_status st = vkjit_var_type(g_ir, v->id, &t);
  if (st != VKJIT_OK) return raise_status(st);
  return PyLong_FromUnsignedLong(t);
}

PyObject* var_steal(PyObject* self, PyObject*) {
  VarObject* v = (VarObject*)self;
  if (!v->live) Py_RETURN_NONE;
  v->live = 0;
  return PyLong_FromUnsignedLong(v->id);
}

PyObject* var_clone(PyObject* self, PyObject*) {  // types.rs:87-92
  VarObject* v = self_live(self);
  if (!v || !ensure_ir()) return nullptr;
  const vkjit_status st = vkjit_inc_ref(g_ir, v->id);
  if (st != VKJIT_OK) return raise_status(st);
  return wrap(v->id);
}

PyObject* var_bop(PyObject* self, PyObject* args) {
  int kind = 0, swap = 0;
  PyObject* rhs = nullptr;
  if (!PyArg_ParseTuple(args, "iO|p", &kind, &rhs, &swap)) return nullptr;
  return swap ? do_bop(kind, rhs, self) : do_bop(kind, self, rhs);
}

#define VK_NAMED(name, kind) \
  PyObject* name(PyObject* self, PyObject* rhs) { return do_bop(kind, self, rhs); }
VK_NAMED(var_lt, VKJIT_BOP_LT)
VK_NAMED(var_gt, VKJIT_BOP_GT)
VK_NAMED(var_eq, VKJIT_BOP_EQ)
VK_NAMED(var_leq, VKJIT_BOP_LEQ)
VK_NAMED(var_geq, VKJIT_BOP_GEQ)
VK_NAMED(var_neq, VKJIT_BOP_NEQ)
#undef VK_NAMED

// Ir::cast / bitcast return the operand's own id when the types already match (internal.rs:283-290), without a new
// reference: the wrapper takes one.
PyObject* cast_like(PyObject* self, PyObject* arg, bool bits) {
  VarObject* v = self_live(self);
  if (!v || !ensure_ir()) return nullptr;
  const unsigned long ty = PyLong_AsUnsignedLong(arg);
  if (ty == (unsigned long)-1 && PyErr_Occurred()) return nullptr;
  vkjit_var out = 0;
  vkjit_status st = bits ? vkjit_bitcast(g_ir, v->id, (vkjit_type)ty, &out) : vkjit_cast(g_ir, v->id, (vkjit_type)ty, &out);
  if (st != VKJIT_OK) return raise_status(st);
  if (out == v->id) {
    st = vkjit_inc_ref(g_ir, out);
    if (st != VKJIT_OK) return raise_status(st);
  }
  return wrap(out);
}
PyObject* var_cast(PyObject* self, PyObject* arg) { return cast_like(self, arg, false); }
PyObject* var_bitcast(PyObject* self, PyObject* arg) { return cast_like(self, arg, true); }

PyObject* var_own(PyObject* cls, PyObject* arg) {  // classmethod: adopt an id that already carries one reference
  const unsigned long id = PyLong_AsUnsignedLong(arg);
  if (id == (unsigned long)-1 && PyErr_Occurred()) return nullptr;
  if (!ensure_ir()) return nullptr;
  return wrap((PyTypeObject*)cls, (vkjit_var)id);
}

PyMethodDef var_methods[] = {
    {"id", var_id, METH_NOARGS, "VarId of this variable (None once consumed)"},
    {"ty", var_ty, METH_NOARGS, "VarType code"},
    {"_steal", var_steal, METH_NOARGS, "give up ownership; returns the id"},
    {"_clone", var_clone, METH_NOARGS, "a second owner of the same variable"},
    {"_bop", var_bop, METH_VARARGS, "_bop(kind, rhs, swap=False)"},
    {"lt", var_lt, METH_O, nullptr}, {"gt", var_gt, METH_O, nullptr}, {"eq", var_eq, METH_O, nullptr},
    {"leq", var_leq, METH_O, nullptr}, {"geq", var_geq, METH_O, nullptr}, {"neq", var_neq, METH_O, nullptr},
    {"cast", var_cast, METH_O, "cast(ty)"},
    {"bitcast", var_bitcast, METH_O, "bitcast(ty)"},
    {"_own", var_own, METH_O | METH_CLASS, "_own(id): wrap an id that already carries one reference"},
    {nullptr, nullptr, 0, nullptr},
};

PyObject* var_get_id(PyObject* self, void*) { return var_id(self, nullptr); }
int var_set_id(PyObject* self, PyObject* value, void*) {  // `*self = ret` of setattr / scatter (vkjit-rust types.rs:152-189)
  VarObject* v = (VarObject*)self;
  if (!value || value == Py_None) { v->live = 0; return 0; }
  const unsigned long id = PyLong_AsUnsignedLong(value);
  if (id == (unsigned long)-1 && PyErr_Occurred()) return -1;
  v->id = (uint32_t)id; v->live = 1; v->gen = g_gen;
  return 0;
}
PyGetSetDef var_getset[] = {
    {"_id", var_get_id, var_set_id, "the owned VarId (None once consumed); assigning transfers ownership", nullptr},
    {nullptr, nullptr, nullptr, nullptr, nullptr},
};

int var_init(PyObject* self, PyObject* args, PyObject* kwds) {  // #[new] __new__(arg) (types.rs:117-120)
  PyObject* arg = nullptr;
  if (kwds && PyDict_GET_SIZE(kwds)) { PyErr_SetString(PyExc_TypeError, "Var() takes no keyword arguments"); return -1; }
  if (!PyArg_ParseTuple(args, "O", &arg)) return -1;
  if (!ensure_ir()) return -1;
  VarObject* v = (VarObject*)self;
  Operand o;
  if (!coerce(arg, o)) return -1;
  if (o.keep) {  // take over the slow path's object
    ((VarObject*)o.keep)->live = 0;
    Py_DECREF(o.keep);
  } else if (!o.temp) {
    const vkjit_status st = vkjit_inc_ref(g_ir, o.id);
    if (st != VKJIT_OK) { raise_status(st); return -1; }
  }
  if (v->live && v->gen == g_gen) vkjit_dec_ref(g_ir, v->id);
  v->id = o.id; v->live = 1; v->gen = g_gen;
  return 0;
}

void var_dealloc(PyObject* self) {  // Drop (types.rs:94-98)
  VarObject* v = (VarObject*)self;
  if (v->live && g_ir && v->gen == g_gen) vkjit_dec_ref(g_ir, v->id);
  v->live = 0;
  Py_TYPE(self)->tp_free(self);
}

PyTypeObject VarBase_Type = {
    PyVarObject_HEAD_INIT(nullptr, 0)
    "vkjit_b200._native.VarBase", /* tp_name */
    sizeof(VarObject),            /* tp_basicsize */
};

// ---- module functions -----------------------------------------------------------------------------------------------
PyObject* mod_bind(PyObject*, PyObject* args) {
  PyObject* owner = nullptr;
  unsigned long long handle = 0;
  if (!PyArg_ParseTuple(args, "OK", &owner, &handle)) return nullptr;
  Py_INCREF(owner);
  Py_XSETREF(g_ir_owner, owner);
  g_ir = (vkjit_ir*)(uintptr_t)handle;
  ++g_gen;
  Py_RETURN_NONE;
}

PyObject* mod_unbind(PyObject*, PyObject*) {
  g_ir = nullptr;
  Py_CLEAR(g_ir_owner);
  Py_RETURN_NONE;
}

PyObject* mod_configure(PyObject*, PyObject* args) {
  PyObject *cls = nullptr, *slow = nullptr, *lazy = nullptr, *errs = nullptr;
  if (!PyArg_ParseTuple(args, "OOOO", &cls, &slow, &lazy, &errs)) return nullptr;
  if (!PyType_Check(cls) || !PyType_IsSubtype((PyTypeObject*)cls, &VarBase_Type)) {
    PyErr_SetString(PyExc_TypeError, "configure: the Var class must derive from VarBase");
    return nullptr;
  }
  if (!PyTuple_Check(errs) || PyTuple_GET_SIZE(errs) != 4) { PyErr_SetString(PyExc_TypeError, "configure: four exception classes"); return nullptr; }
  Py_INCREF(cls);
  { PyTypeObject* old = g_var_cls; g_var_cls = (PyTypeObject*)cls; Py_XDECREF(old); }
  Py_INCREF(slow); Py_XSETREF(g_slow_coerce, slow);
  Py_INCREF(lazy); Py_XSETREF(g_lazy_bind, lazy);
  for (int i = 0; i < 4; ++i) { PyObject* e = PyTuple_GET_ITEM(errs, i); Py_INCREF(e); Py_XSETREF(g_errs[i], e); }
  Py_RETURN_NONE;
}

PyObject* mod_coerce(PyObject*, PyObject* v) {
  if (!ensure_ir()) return nullptr;
  return coerce_to_object(v);
}

PyObject* mod_bop(PyObject*, PyObject* args) {
  int kind = 0;
  PyObject *a = nullptr, *b = nullptr;
  if (!PyArg_ParseTuple(args, "iOO", &kind, &a, &b)) return nullptr;
  return do_bop(kind, a, b);
}

PyObject* mod_uop(PyObject*, PyObject* args) {
  int kind = 0;
  PyObject* a = nullptr;
  if (!PyArg_ParseTuple(args, "iO", &kind, &a)) return nullptr;
  return do_uop(kind, a);
}

PyObject* mod_select(PyObject*, PyObject* args) {  // vkjit-rust functions.rs:24-31
  PyObject *c = nullptr, *a = nullptr, *b = nullptr;
  if (!PyArg_ParseTuple(args, "OOO", &c, &a, &b)) return nullptr;
  if (!ensure_ir()) return nullptr;
  Operand oc, oa, ob;
  if (!coerce(c, oc)) return nullptr;
  if (!coerce(a, oa)) { release(oc); return nullptr; }
  if (!coerce(b, ob)) { release(oc); release(oa); return nullptr; }
  vkjit_var out = 0;
  const vkjit_status st = vkjit_select(g_ir, oc.id, oa.id, ob.id, &out);
  release(oc); release(oa); release(ob);
  if (st != VKJIT_OK) return raise_status(st);
  return wrap(out);
}

PyObject* mod_arange(PyObject*, PyObject* args) {  // vkjit-rust functions.rs:9-11
  unsigned int ty = 0;
  unsigned long long n = 0;
  if (!PyArg_ParseTuple(args, "IK", &ty, &n)) return nullptr;
  if (!ensure_ir()) return nullptr;
  vkjit_var out = 0;
  const vkjit_status st = vkjit_arange(g_ir, ty, (size_t)n, &out);
  if (st != VKJIT_OK) return raise_status(st);
  return wrap(out);
}

// ids of a sequence of Vars (eval(schedule: &PyList), functions.rs:9-23)
bool collect_ids(PyObject* seq, std::vector<vkjit_var>& ids) {
  PyObject* fast = PySequence_Fast(seq, "expected a sequence of Var");
  if (!fast) return false;
  const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
  ids.reserve((size_t)n);
  for (Py_ssize_t i = 0; i < n; ++i) {
    PyObject* it = PySequence_Fast_GET_ITEM(fast, i);
    if (!PyObject_TypeCheck(it, &VarBase_Type) || !((VarObject*)it)->live || ((VarObject*)it)->gen != g_gen) {
      Py_DECREF(fast);
      PyErr_SetString(PyExc_TypeError, "expected a sequence of Var");
      return false;
    }
    ids.push_back(((VarObject*)it)->id);
  }
  Py_DECREF(fast);
  return true;
}

PyObject* eval_like(PyObject* seq, bool run) {
  if (!ensure_ir()) return nullptr;
  std::vector<vkjit_var> ids;
  if (!collect_ids(seq, ids)) return nullptr;
  vkjit_status st;
  Py_BEGIN_ALLOW_THREADS  // a cache miss runs NVRTC for tens of milliseconds
  st = run ? vkjit_eval(g_ir, ids.data(), ids.size()) : vkjit_schedule(g_ir, ids.data(), ids.size());
  Py_END_ALLOW_THREADS
  if (st != VKJIT_OK) return raise_status(st);
  Py_RETURN_NONE;
}
PyObject* mod_eval(PyObject*, PyObject* seq) { return eval_like(seq, true); }
PyObject* mod_schedule(PyObject*, PyObject* seq) { return eval_like(seq, false); }

PyMethodDef module_methods[] = {
    {"bind", mod_bind, METH_VARARGS, "bind(owner, handle): the global Ir the front-end records into"},
    {"unbind", mod_unbind, METH_NOARGS, "forget the global Ir (Vars still alive are no longer released)"},
    {"configure", mod_configure, METH_VARARGS, "configure(VarClass, slow_coerce, lazy_bind, (errors...))"},
    {"coerce", mod_coerce, METH_O, "TryFrom<&PyAny> for Var (types.rs:47-82)"},
    {"bop", mod_bop, METH_VARARGS, "bop(kind, lhs, rhs)"},
    {"uop", mod_uop, METH_VARARGS, "uop(kind, x)"},
    {"select", mod_select, METH_VARARGS, "select(condition, x, y)"},
    {"arange", mod_arange, METH_VARARGS, "arange(ty, n)"},
    {"eval", mod_eval, METH_O, "eval([vars])"},
    {"schedule", mod_schedule, METH_O, "schedule([vars])"},
    {nullptr, nullptr, 0, nullptr},
};

PyModuleDef module_def = {PyModuleDef_HEAD_INIT, "_native", "native half of the vkjit Python front-end (see vkjit_b200/vkjit.py)", -1,
                          module_methods, nullptr, nullptr, nullptr, nullptr};

}  // namespace

PyMODINIT_FUNC PyInit__native(void) {
  VarBase_Type.tp_flags = Py_TPFLAGS_DEFAULT | Py_TPFLAGS_BASETYPE;
  VarBase_Type.tp_doc = "`#[pyclass] pub struct Var(VarId)` (vkjit-python/src/types.rs:84-85): owns one reference count";
  VarBase_Type.tp_new = PyType_GenericNew;
  VarBase_Type.tp_init = var_init;
  VarBase_Type.tp_dealloc = var_dealloc;
  VarBase_Type.tp_as_number = &var_as_number;
  VarBase_Type.tp_richcompare = var_richcompare;
  VarBase_Type.tp_hash = PyBaseObject_Type.tp_hash;  // identity, as `==` stays identity
  VarBase_Type.tp_methods = var_methods;
  VarBase_Type.tp_getset = var_getset;
  if (PyType_Ready(&VarBase_Type) < 0) return nullptr;
  PyObject* m = PyModule_Create(&module_def);
  if (!m) return nullptr;
  Py_INCREF(&VarBase_Type);
  if (PyModule_AddObject(m, "VarBase", (PyObject*)&VarBase_Type) < 0) { Py_DECREF(&VarBase_Type); Py_DECREF(m); return nullptr; }
  return m;
}
