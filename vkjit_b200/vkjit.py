"""Drop-in mirror of the vkjit-python module (reference libs/vkjit-python/src/{lib,types,functions}.rs).

Reference surface (lib.rs:18-26): class `Var`, functions `eval`, `var`, `ir`, `linspace`.  Same
names, same argument coercion order (types.rs:50-81), same lazy `__str__` (types.rs:148-154),
on top of one global `Ir` (lib.rs:14-16) — here the B200 backend's `Ir` through the C ABI.

Additions the reference lacks (needed by the Monte-Carlo config, which is
driven from Python): `arange`, comparisons, `select`, `gather`/`scatter`/`scatter_add`, bit ops,
unary math, reductions, `prefix_sum`, `compress`, `numpy()`/`tolist()` for every dtype and
`__truediv__` (the reference only defines the Python-2 `__div__`, types.rs:136-139).
The known reference defect that `eval()` calls an unexported `.id()` (functions.rs:15 vs
types.rs:99-102) is fixed: `Var.id()` exists.
"""
from __future__ import annotations

import threading

import numpy as np

from . import ir as _ir
from ._capi import VkjitError, VkjitNoDeviceError, VkjitSizeError, VkjitTypeError
from .ir import Bop, Red, Uop, VarType

try:
    from . import _native
except ImportError as e:  # pragma: no cover - build error
    raise ImportError(
        "vkjit_b200/_native*.so is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'). "
        "The Python front-end is a compiled module, as the reference's (pyo3); there is no pure-Python path.") from e

_lock = threading.RLock()
_IR = None


class _GlobalIr(_ir.Ir):
    """The process-wide Ir.  Closing it first detaches the native Var type, so handles that outlive it turn into
    no-ops on release instead of touching a destroyed Ir."""

    def close(self):
        global _IR
        if self._h is not None and self._h.value:
            _native.unbind()
            with _lock:
                if _IR is self:
                    _IR = None
        super().close()


def _global_ir() -> _ir.Ir:
    """`lazy_static! { pub static ref IR: Mutex<Ir> }` (lib.rs:14-16)"""
    global _IR
    with _lock:
        if _IR is None:
            # `Ir::new()` creates the backend (internal.rs:168-182 -> VulkanBackend::create); here the device is
            # bound on first use.  Without a B200 traces can still be built and printed; uploads and eval then
            # fail loudly with VkjitNoDeviceError (there is no CPU path).
            try:
                _ir.init(-1)
            except VkjitNoDeviceError:
                pass
            _IR = _GlobalIr()
            _native.bind(_IR, _IR._h.value)  # the native Var records into this Ir (and keeps it alive)
        return _IR


def _is_int(x):
    return isinstance(x, (int, np.integer)) and not isinstance(x, (bool, np.bool_))


def _coerce_slow(value) -> "Var":
    """The part of `TryFrom<&PyAny> for Var` (types.rs:47-82) that needs NumPy: NumPy scalars, arrays and
    sequences ([u32], [i32], [f32] in that order).  Var / bool / int / float are handled natively
    (csrc/pyfront.cpp: coerce)."""
    g = _global_ir()
    if isinstance(value, (bool, np.bool_)):
        # pyo3 extracts a Python bool as u32 first (bool is an int): True -> UInt32(1)
        return Var._own(g.const_u32(int(value)))
    if _is_int(value):
        v = int(value)
        if 0 <= v <= 0xFFFFFFFF:
            return Var._own(g.const_u32(v))
        if -2 ** 31 <= v < 0:
            return Var._own(g.const_i32(v))
        raise TypeError("Not a valid argument!")
    if isinstance(value, (float, np.floating)):
        return Var._own(g.const_f32(float(value)))
    if isinstance(value, np.ndarray):
        if value.dtype == np.uint32:
            return Var._own(g.array_u32(value))
        if value.dtype == np.int32:
            return Var._own(g.array_i32(value))
        if value.dtype == np.float32:
            return Var._own(g.array_f32(value))
        if value.dtype == np.bool_:
            return Var._own(g.array_bool(value))
        value = value.tolist()
    if isinstance(value, (list, tuple)):
        items = list(value)
        if all(_is_int(x) or isinstance(x, (bool, np.bool_)) for x in items):
            if all(0 <= int(x) <= 0xFFFFFFFF for x in items):
                return Var._own(g.array_u32(np.asarray(items, dtype=np.uint32)))
            if all(-2 ** 31 <= int(x) < 2 ** 31 for x in items):
                return Var._own(g.array_i32(np.asarray(items, dtype=np.int32)))
            raise TypeError("Not a valid argument!")
        if all(isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, bool) for x in items):
            return Var._own(g.array_f32(np.asarray(items, dtype=np.float32)))
    raise TypeError("Not a valid argument!")


_coerce = _native.coerce   # any accepted value -> a new owned Var (a Var argument is cloned)


class Var(_native.VarBase):
    """`#[pyclass] pub struct Var(VarId)` (types.rs:84-85).  Owns exactly one reference count:
    cloning bumps it (types.rs:87-92), dropping releases it (types.rs:94-98).

    The base type is native (csrc/pyfront.cpp): construction from a value (`#[new]`, types.rs:117-120), `id()`,
    `ty()`, the arithmetic / bit / comparison operators (types.rs:124-139) with their reflected forms, the named
    comparisons `lt gt eq leq geq neq`, `cast`, `bitcast`, and the ownership helpers `_own`, `_steal`, `_clone`,
    `_id`.  This subclass adds what is not on the trace-building path."""

    __slots__ = ()
    __array_ufunc__ = None   # NumPy operands defer to Var's reflected operators instead of broadcasting over it

    def __div__(self, rhs): return self._bop(Bop.Div, rhs)        # types.rs:136-139 (Python-2 name kept)

    def __repr__(self):  # types.rs:140-147
        g = _global_ir()
        if g.is_buffer(self._id):
            return f"array(dtype = {VarType.name(g.ty(self._id))}, {g.str(self._id)})"
        return g.str(self._id)

    def __str__(self):  # types.rs:148-154: evaluates on demand
        g = _global_ir()
        if not g.is_buffer(self._id):
            g.eval([self._id])
        return g.str(self._id)

    # -- additions (N2) ------------------------------------------------------------
    def tolist(self):
        """types.rs:121-123 reads f32 only; here every dtype is readable (evaluates if needed)."""
        return self.numpy().tolist()

    def then_else(self, then, other) -> "Var":  # vkjit-rust types.rs:160-168
        return select(self, then, other)

    def get(self, idx) -> "Var":  # vkjit-rust types.rs:190-192
        return gather(self, idx)

    def getattr(self, idx: int) -> "Var":  # vkjit-rust types.rs:149-151
        return Var._own(_global_ir().getattr(self._id, idx))

    def setattr(self, var, idx: int):  # vkjit-rust types.rs:152-159: `*self = ret`
        v = _coerce(var)
        g = _global_ir()
        new = g.setattr(self._id, v._id, idx)
        g.dec_ref_count(self._id)
        self._id = new

    def scatter_with(self, to: "Var", idx, condition=None):  # vkjit-rust types.rs:172-189
        self.scatter(to, idx, condition)

    def to_vec(self):  # vkjit-rust types.rs:193-196 (as_slice: the var must be evaluated)
        g = _global_ir()
        return g.as_slice(self._id, g.ty(self._id)).tolist()

    def scatter(self, to: "Var", idx, condition=None):  # vkjit-rust types.rs:169-189
        i = _coerce(idx)
        c = None if condition is None else _coerce(condition)
        g = _global_ir()
        new = g.scatter(self._id, to._id, i._id, None if c is None else c._id)
        g.dec_ref_count(self._id)
        self._id = new

    def scatter_add(self, to: "Var", idx, condition=None):
        i = _coerce(idx)
        c = None if condition is None else _coerce(condition)
        g = _global_ir()
        new = g.scatter_add(self._id, to._id, i._id, None if c is None else c._id)
        g.dec_ref_count(self._id)
        self._id = new

    def numpy(self) -> np.ndarray:
        g = _global_ir()
        if not g.is_buffer(self._id):
            g.eval([self._id])
        return g.to_numpy(self._id)

    @property
    def __cuda_array_interface__(self):
        """Zero-copy export (CUDA Array Interface v3) of the evaluated device array; consumers must order their
        work after the backend stream (`vkjit_b200.stream_ptr()`), which is what `stream` advertises."""
        g = _global_ir()
        if not g.is_buffer(self._id):
            g.eval([self._id])
        ty = g.ty(self._id)
        typestr = {VarType.F32: "<f4", VarType.U32: "<u4", VarType.I32: "<i4", VarType.Bool: "<u4"}[ty]
        return {"shape": (g.size(self._id),), "typestr": typestr, "data": (g.device_ptr(self._id), False), "version": 3,
                "strides": None, "stream": _ir.stream_ptr() or None}

    def __dlpack__(self, *, stream=None, max_version=None, dl_device=None, copy=None):
        """DLPack export (zero-copy; evaluates first; the backend stream is synchronised, so any consumer stream
        is safe).  `torch.from_dlpack(v)` / `cupy.from_dlpack(v)` work on a Var."""
        if copy:
            raise BufferError("vkjit_b200 exports views only")
        g = _global_ir()
        if not g.is_buffer(self._id):
            g.eval([self._id])
        return g.to_dlpack(self._id)

    def __dlpack_device__(self):
        return (2, _ir.device_index())   # kDLCUDA

    def sum(self): return Var._own(_global_ir().reduce(Red.Sum, self._id))
    def min(self): return Var._own(_global_ir().reduce(Red.Min, self._id))
    def max(self): return Var._own(_global_ir().reduce(Red.Max, self._id))
    def prefix_sum(self, exclusive=True): return Var._own(_global_ir().prefix_sum(self._id, exclusive))

    def compress(self):
        """Indices of the set lanes of a Bool var (stable), and their count."""
        out, n = _global_ir().compress(self._id)
        return Var._own(out), n


# -- module functions (functions.rs:9-52) ---------------------------------------------------
def eval(schedule):  # noqa: A001 - the reference's name
    """`eval(schedule: &PyList)` (functions.rs:9-23)"""
    _native.eval(schedule)


def var(*args) -> Var:
    """`var(*args)` (functions.rs:25-34): one argument is coerced as is, several form a sequence."""
    return _coerce(args[0]) if len(args) == 1 else _coerce(list(args))


def ir() -> str:
    """`format!("{:#?}", IR)` (functions.rs:36-39)"""
    return _global_ir().repr()


def linspace(start, stop, num: int) -> Var:
    """functions.rs:41-52 — asserts start.ty() == stop.ty(); endpoint excluded."""
    a, b = _coerce(start), _coerce(stop)
    assert a.ty() == b.ty()
    return Var._own(_global_ir().linspace(a.ty(), a._id, b._id, num))


# -- the Rust front-end's free functions (vkjit-rust/src/functions.rs) -------------------------
def schedule(vars_):
    """`schedule!(a, b, ...)` / schedule_internal (functions.rs:72-82): queue vars for the next eval."""
    _native.schedule(vars_)


def repr_ir() -> str:  # functions.rs:54-56
    return _global_ir().repr()


def struct(*members) -> Var:
    """`Var::from(&[Var])` (vkjit-rust types.rs:99-105): a struct-typed var from its members."""
    vs = [_coerce(m) for m in members]
    return Var._own(_global_ir().struct_init([v._id for v in vs]))


def gather_with(src: Var, idx, condition=None) -> Var:  # functions.rs:39-52
    return gather(src, idx, condition)


def ones(ty: int) -> Var:  # Ir::ones (internal.rs:265-282; not re-exported by the reference front-ends)
    return Var._own(_global_ir().ones(ty))


# -- additions -------------------------------------------------------------------------------
def arange(ty: int, num: int) -> Var:  # vkjit-rust functions.rs:9-11
    return _native.arange(ty, num)


def zeros(ty: int) -> Var:  # vkjit-rust functions.rs:5-7
    return Var._own(_global_ir().zeros(ty))


def select(condition, x, y) -> Var:  # vkjit-rust functions.rs:24-31
    return _native.select(condition, x, y)


def gather(src: Var, idx, condition=None) -> Var:  # vkjit-rust functions.rs:32-52
    i = _coerce(idx)
    c = None if condition is None else _coerce(condition)
    return Var._own(_global_ir().gather(src._id, i._id, None if c is None else c._id))


def _u(kind):
    def f(x):
        return _native.uop(kind, x)
    return f


sqrt, exp, log, sin, cos = _u(Uop.Sqrt), _u(Uop.Exp), _u(Uop.Log), _u(Uop.Sin), _u(Uop.Cos)


def minimum(a, b) -> Var:
    return _native.bop(Bop.Min, a, b)


def maximum(a, b) -> Var:
    return _native.bop(Bop.Max, a, b)


def compress(values: Var, mask: Var):
    out, n = _global_ir().compress_values(values._id, mask._id)
    return Var._own(out), n


def from_cuda_array(obj) -> Var:
    """Zero-copy import of any object exposing `__cuda_array_interface__` (torch tensors, CuPy arrays): 1-D,
    contiguous, f32/u32/i32.  The object must stay alive while the returned Var is used, and its producer must be
    ordered before the backend stream."""
    cai = obj.__cuda_array_interface__
    if len(cai["shape"]) != 1 or cai.get("strides") not in (None, (4,)):
        raise TypeError("Not a valid argument!")
    ty = {"<f4": VarType.F32, "<u4": VarType.U32, "<i4": VarType.I32}.get(cai["typestr"])
    if ty is None:
        raise TypeError("Not a valid argument!")
    return Var._own(_global_ir().array_wrap_device(ty, int(cai["data"][0]), int(cai["shape"][0])))


def from_dlpack(obj) -> Var:
    """Zero-copy import of a DLPack capsule or of any object with `__dlpack__` (torch / CuPy / JAX arrays): 1-D,
    contiguous, f32/u32/i32, on this GPU.  The tensor is kept alive by the returned Var's array and released
    when that array is dropped; its producer must be ordered before the backend stream."""
    cap = obj.__dlpack__() if hasattr(obj, "__dlpack__") else obj
    return Var._own(_global_ir().from_dlpack(cap))


def sync():
    _ir.sync()


_native.configure(Var, _coerce_slow, _global_ir, (VkjitError, VkjitTypeError, VkjitSizeError, VkjitNoDeviceError))
