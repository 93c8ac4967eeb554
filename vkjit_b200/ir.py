"""Host-side mirror of `vkjit_core::Ir` (libs/vkjit-core/src/internal.rs:126-542).

Same method names, argument order and error behaviour as the reference's `Ir`, so
the parity tests read like libs/vkjit-core/src/test.rs.  Every method is a thin
call through the C ABI (include/vkjit_b200.h); all work happens in the native
library.  Var ids are plain ints (`VarId`).
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Sequence

import numpy as np

from ._capi import CApi, Stats, product_api


class VarType:
    """vartype.rs:24-33 — codes as in include/vkjit_b200.h; struct types are interned per Ir."""
    Void, Bool, U32, I32, F32 = 1, 2, 3, 4, 5
    STRUCT_BASE = 16
    _names = {1: "Void", 2: "Bool", 3: "U32", 4: "I32", 5: "F32"}
    _np = {2: np.uint32, 3: np.uint32, 4: np.int32, 5: np.float32}

    @classmethod
    def name(cls, ty: int) -> str:
        return cls._names.get(ty, f"Struct#{ty - 16}")

    @classmethod
    def numpy(cls, ty: int):
        return cls._np[ty]


class Bop:
    Add, Sub, Mul, Div, Lt, Gt, Eq, Leq, Geq, Neq = range(10)
    And, Or, Xor, Shl, Shr, Min, Max = range(16, 23)


class Uop:
    Neg, Abs, Not, Sqrt, Exp, Log, Sin, Cos = range(8)


class Red:
    Sum, Min, Max = range(3)


def _u32arr(ids: Sequence[int]):
    return (C.c_uint32 * len(ids))(*[int(i) for i in ids])


# ---- pinned result arrays ------------------------------------------------------------------------------------------
_PINNED_MIN_BYTES = 64 << 10
_PINNED_KEEP_BYTES = 1 << 30      # blocks kept for reuse, in total
_pinned_free = {}                 # capacity -> [ptr]
_pinned_kept = 0


def _pinned_release(api, ptr, cap):
    global _pinned_kept
    if _pinned_kept + cap <= _PINNED_KEEP_BYTES and len(_pinned_free.setdefault(cap, [])) < 8:
        _pinned_free[cap].append(ptr)
        _pinned_kept += cap
        return
    try:
        api._fn["host_free"](C.c_void_p(ptr))    # status ignored: at interpreter exit the context may be gone
    except Exception:
        pass


def _pinned_empty(api, n, dtype):
    """np.ndarray of n elements in pinned host memory (only with the product library and an initialised backend;
    otherwise an ordinary array).  The block goes back to the pool when the last view of the array dies."""
    global _pinned_kept
    import weakref
    nbytes = n * np.dtype(dtype).itemsize
    if not api.has("host_alloc") or api.prefix != "vkjit_":
        return np.empty(n, dtype=dtype)
    cap = 1 << max(16, (nbytes - 1).bit_length())       # power-of-two size classes
    lst = _pinned_free.get(cap)
    if lst:
        ptr = lst.pop()
        _pinned_kept -= cap
    else:
        p = C.c_void_p()
        try:
            api.call("host_alloc", cap, C.byref(p))
        except Exception:
            return np.empty(n, dtype=dtype)
        ptr = p.value
    buf = (C.c_char * cap).from_address(ptr)
    weakref.finalize(buf, _pinned_release, api, ptr, cap)
    return np.frombuffer(buf, dtype=dtype, count=n)


def pinned_empty(n: int, dtype=np.float32) -> np.ndarray:
    """An empty array in pinned host memory: build inputs here and `array_*` uploads them with one direct DMA."""
    from ._capi import product_api
    return _pinned_empty(product_api(), n, dtype)


class Ir:
    def __init__(self, _api: Optional[CApi] = None):
        self.api = _api or product_api()
        h = C.c_void_p()
        self.api.call("ir_create", C.byref(h))
        self._h = h

    def close(self):
        if self._h is not None and self._h.value:
            self.api.call("ir_destroy", self._h)
        self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers ---------------------------------------------------------
    def _new(self, name, *args) -> int:
        out = C.c_uint32()
        self.api.call(name, self._h, *args, C.byref(out))
        return out.value

    def struct_type(self, elems: Sequence[int]) -> int:
        return self._new("type_struct", _u32arr(elems), len(elems))

    def struct_type_elems(self, ty: int):
        n = C.c_size_t()
        self.api.call("type_struct_len", self._h, ty, C.byref(n))
        out = []
        for i in range(n.value):
            e = C.c_uint32()
            self.api.call("type_struct_elem", self._h, ty, i, C.byref(e))
            out.append(e.value)
        return out

    # -- constructors (internal.rs:229-400) -------------------------------
    def const_f32(self, v: float) -> int:
        return self._new("const_f32", C.c_float(v))

    def const_i32(self, v: int) -> int:
        return self._new("const_i32", C.c_int32(v))

    def const_u32(self, v: int) -> int:
        return self._new("const_u32", C.c_uint32(v))

    def const_bool(self, v: bool) -> int:
        return self._new("const_bool", C.c_int32(1 if v else 0))

    def _array(self, name, data, dtype) -> int:
        a = np.ascontiguousarray(data, dtype=dtype)
        return self._new(name, a.ctypes.data_as(C.c_void_p), a.size)

    def array_f32(self, data) -> int:
        return self._array("array_f32", data, np.float32)

    def array_i32(self, data) -> int:
        return self._array("array_i32", data, np.int32)

    def array_u32(self, data) -> int:
        return self._array("array_u32", data, np.uint32)

    def array_bool(self, data) -> int:
        return self._array("array_bool", np.asarray(data).astype(np.uint32), np.uint32)

    def array_empty(self, ty: int, n: int) -> int:
        return self._new("array_empty", ty, n)

    def array_wrap_device(self, ty: int, device_ptr: int, n: int) -> int:
        """Zero-copy view of foreign device memory (not owned)."""
        return self._new("array_wrap_device", ty, C.c_uint64(device_ptr), n)

    # ---- DLPack: capsules named "dltensor" holding a DLManagedTensor*
    def to_dlpack(self, id: int):
        """Zero-copy export of an evaluated var as a DLPack capsule (the capsule holds one reference on the var)."""
        mt = C.c_void_p()
        self.api.call("var_to_dlpack", self._h, id, C.byref(mt))
        return _dl_capsule_new(self.api, mt.value)

    def from_dlpack(self, capsule) -> int:
        """Zero-copy import of a DLPack capsule (1-D contiguous f32/i32/u32 CUDA tensor); the capsule is consumed."""
        py = C.pythonapi
        if not py.PyCapsule_IsValid(C.py_object(capsule), b"dltensor"):
            raise TypeError("Not a valid argument!")   # not a capsule, or already consumed
        mt = py.PyCapsule_GetPointer(C.py_object(capsule), b"dltensor")
        out = C.c_uint32()
        self.api.call("var_from_dlpack", self._h, C.c_void_p(mt), C.byref(out))   # raises: the capsule stays valid
        py.PyCapsule_SetName(C.py_object(capsule), b"used_dltensor")              # ownership moved to the library
        return out.value

    def arange(self, ty: int, num: int) -> int:
        return self._new("arange", ty, num)

    def linspace(self, ty: int, start: int, stop: int, num: int) -> int:
        return self._new("linspace", ty, start, stop, num)

    def zeros(self, ty: int) -> int:
        return self._new("zeros", ty)

    def ones(self, ty: int) -> int:
        return self._new("ones", ty)

    def cast(self, src: int, ty: int) -> int:
        return self._new("cast", src, ty)

    def bop(self, kind: int, lhs: int, rhs: int) -> int:
        return self._new("bop", kind, lhs, rhs)

    def uop(self, kind: int, src: int) -> int:
        return self._new("uop", kind, src)

    def bitcast(self, src: int, ty: int) -> int:
        return self._new("bitcast", src, ty)

    def select(self, cond: int, lhs: int, rhs: int) -> int:
        return self._new("select", cond, lhs, rhs)

    def struct_init(self, elems: Sequence[int]) -> int:
        return self._new("struct_init", _u32arr(elems), len(elems))

    def getattr(self, src: int, idx: int) -> int:
        return self._new("getattr", src, idx)

    def setattr(self, dst: int, src: int, idx: int) -> int:
        return self._new("setattr", dst, src, idx)

    def gather(self, src: int, idx: int, active: Optional[int] = None) -> int:
        return self._new("gather", src, idx, 0 if active is None else 1, 0 if active is None else active)

    def scatter(self, src: int, dst: int, idx: int, active: Optional[int] = None) -> int:
        return self._new("scatter", src, dst, idx, 0 if active is None else 1, 0 if active is None else active)

    def scatter_add(self, src: int, dst: int, idx: int, active: Optional[int] = None) -> int:
        return self._new("scatter_add", src, dst, idx, 0 if active is None else 1, 0 if active is None else active)

    # extension unary helpers
    def neg(self, a): return self.uop(Uop.Neg, a)
    def abs(self, a): return self.uop(Uop.Abs, a)
    def not_(self, a): return self.uop(Uop.Not, a)
    def sqrt(self, a): return self.uop(Uop.Sqrt, a)
    def exp(self, a): return self.uop(Uop.Exp, a)
    def log(self, a): return self.uop(Uop.Log, a)
    def sin(self, a): return self.uop(Uop.Sin, a)
    def cos(self, a): return self.uop(Uop.Cos, a)

    # -- introspection / lifetime ------------------------------------------
    def ty(self, id: int) -> int:
        return self._new("var_type", id)

    def ref_count(self, id: int) -> int:
        return self._new("var_ref_count", id)

    def num_vars(self) -> int:
        n = C.c_size_t()
        self.api.call("var_count", self._h, C.byref(n))
        return n.value

    def num_arrays(self) -> int:
        n = C.c_size_t()
        self.api.call("array_count", self._h, C.byref(n))
        return n.value

    def is_buffer(self, id: int) -> bool:
        o = C.c_int32()
        self.api.call("is_buffer", self._h, id, C.byref(o))
        return bool(o.value)

    def size(self, id: int) -> int:
        n = C.c_size_t()
        self.api.call("var_size", self._h, id, C.byref(n))
        return n.value

    def device_ptr(self, id: int) -> int:
        p = C.c_uint64()
        self.api.call("var_device_ptr", self._h, id, C.byref(p))
        return p.value

    def inc_ref_count(self, id: int):
        self.api.call("inc_ref", self._h, id)

    def dec_ref_count(self, id: int):
        self.api.call("dec_ref", self._h, id)

    def _string(self, name, *args) -> str:
        n = C.c_size_t()
        self.api.call(name, self._h, *args, None, 0, C.byref(n))
        buf = C.create_string_buffer(n.value + 1)
        self.api.call(name, self._h, *args, buf, n.value + 1, C.byref(n))
        return buf.value.decode("utf-8", "replace")

    def repr(self) -> str:
        """`format!("{:#?}", ir)`"""
        return self._string("ir_repr")

    def str(self, id: int) -> str:
        """Ir::str for buffers (internal.rs:404-422), `{:?}` of the Var otherwise."""
        return self._string("var_repr", id)

    # -- execute -----------------------------------------------------------
    def schedule(self, ids: Iterable[int]):
        ids = list(ids)
        self.api.call("schedule", self._h, _u32arr(ids), len(ids))

    def eval(self, ids: Iterable[int]):
        ids = list(ids)
        self.api.call("eval", self._h, _u32arr(ids), len(ids))

    def as_slice(self, id: int, ty: int) -> np.ndarray:
        """Ir::as_slice::<T> (internal.rs:443-449): `ty` plays the role of T.  The reference returns a zero-copy view of
        host-visible memory (vulkan/mod.rs:29-34); here the result array of a large read lives in PINNED host memory
        (a pool of vkjit_host_alloc blocks, returned when the array is garbage collected), so the D2H copy is one DMA
        at PCIe speed straight into the array the caller gets — no staging copy, no second pass over the data."""
        n = self.size(id)
        dt = VarType.numpy(ty)
        out = _pinned_empty(self.api, n, dt) if n * 4 >= _PINNED_MIN_BYTES else np.empty(n, dtype=dt)
        self.api.call("read", self._h, id, ty, out.ctypes.data_as(C.c_void_p), out.nbytes)
        return out

    def as_slice_eval(self, id: int, ty: int) -> np.ndarray:
        """eval([id]) if needed, then as_slice."""
        if not self.is_buffer(id):
            self.eval([id])
        return self.as_slice(id, ty)

    def read_into(self, id: int, ty: int, ptr: int, nbytes: int):
        self.api.call("read", self._h, id, ty, C.c_void_p(ptr), nbytes)

    def to_numpy(self, id: int) -> np.ndarray:
        ty = self.ty(id)
        a = self.as_slice(id, ty)
        return a.astype(bool) if ty == VarType.Bool else a

    # -- runtime primitives (extensions) -------------------------------------
    def reduce(self, red: int, id: int) -> int:
        return self._new("reduce", red, id)

    def sum(self, id): return self.reduce(Red.Sum, id)
    def min(self, id): return self.reduce(Red.Min, id)
    def max(self, id): return self.reduce(Red.Max, id)

    def prefix_sum(self, id: int, exclusive: bool = True) -> int:
        return self._new("prefix_sum", id, 1 if exclusive else 0)

    def compress(self, mask: int):
        out, n = C.c_uint32(), C.c_size_t()
        self.api.call("compress", self._h, mask, C.byref(out), C.byref(n))
        return out.value, n.value

    def compress_values(self, values: int, mask: int):
        out, n = C.c_uint32(), C.c_size_t()
        self.api.call("compress_values", self._h, values, mask, C.byref(out), C.byref(n))
        return out.value, n.value

    # -- multi-GPU (product only) ----------------------------------------------
    def arange_sharded(self, ty: int, n: int) -> int:
        return self._new("arange_sharded", ty, n)

    def array_sharded(self, ty: int, data) -> int:
        a = np.ascontiguousarray(data, dtype=VarType.numpy(ty))
        return self._new("array_sharded", ty, a.ctypes.data_as(C.c_void_p), a.size)

    def array_shard_local(self, ty: int, data=None, ptr: int = 0, n: int = 0) -> int:
        """This rank's slice given directly (numpy array, or raw host pointer + element count)."""
        if data is not None:
            a = np.ascontiguousarray(data, dtype=VarType.numpy(ty))
            return self._new("array_shard_local", ty, a.ctypes.data_as(C.c_void_p), a.size)
        return self._new("array_shard_local", ty, C.c_void_p(ptr), n)

    def is_sharded(self, id: int) -> bool:
        o = C.c_int32()
        self.api.call("var_is_sharded", self._h, id, C.byref(o))
        return bool(o.value)

    def shard_base(self, id: int) -> int:
        """Global index of the first element this rank holds (sharded aranges / arrays / compress results)."""
        o = C.c_uint64()
        self.api.call("var_shard_base", self._h, id, C.byref(o))
        return int(o.value)

    def debug_codegen(self, ids: Sequence[int], compile: bool = False, privatize: bool = False):
        ids = list(ids)
        n, cub = C.c_size_t(), C.c_size_t()
        self.api.call("debug_codegen", self._h, _u32arr(ids), len(ids), 2 if privatize else 0, None, 0, C.byref(n), C.byref(cub))
        buf = C.create_string_buffer(n.value + 1)
        self.api.call("debug_codegen", self._h, _u32arr(ids), len(ids), (1 if compile else 0) | (2 if privatize else 0), buf,
                      n.value + 1, C.byref(n), C.byref(cub))
        return buf.value.decode(), cub.value


C.pythonapi.PyCapsule_New.restype = C.py_object
C.pythonapi.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
C.pythonapi.PyCapsule_IsValid.restype = C.c_int
C.pythonapi.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
C.pythonapi.PyCapsule_GetPointer.restype = C.c_void_p
C.pythonapi.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
C.pythonapi.PyCapsule_SetName.restype = C.c_int
C.pythonapi.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
_DL_DESTRUCTOR = C.CFUNCTYPE(None, C.c_void_p)
_dl_destructors = {}


def _dl_capsule_new(api, managed_ptr: int):
    """PyCapsule("dltensor") whose destructor runs the tensor's deleter unless a consumer renamed (= took) it."""
    key = id(api)
    if key not in _dl_destructors:
        delete = getattr(api.lib, "vkjit_dlpack_delete")
        delete.argtypes, delete.restype = [C.c_void_p], None

        def destroy(capsule_ptr):
            cap = C.cast(capsule_ptr, C.py_object)
            if C.pythonapi.PyCapsule_IsValid(cap, b"dltensor"):
                delete(C.pythonapi.PyCapsule_GetPointer(cap, b"dltensor"))
        _dl_destructors[key] = _DL_DESTRUCTOR(destroy)
    return C.pythonapi.PyCapsule_New(managed_ptr, b"dltensor", C.cast(_dl_destructors[key], C.c_void_p))


def _debug_codegen_reduce(self, id: int, red: int, compile: bool = False):
    n, cub = C.c_size_t(), C.c_size_t()
    self.api.call("debug_codegen_reduce", self._h, id, red, 0, None, 0, C.byref(n), C.byref(cub))
    buf = C.create_string_buffer(n.value + 1)
    self.api.call("debug_codegen_reduce", self._h, id, red, 1 if compile else 0, buf, n.value + 1, C.byref(n), C.byref(cub))
    return buf.value.decode(), cub.value


Ir.debug_codegen_reduce = _debug_codegen_reduce


def _debug_codegen_scan(self, ids, mode: int, compile: bool = False):
    """Source (and cubin size) of the fused trace -> scan kernel; mode 0/1 prefix sums, 2 compress, 3 compress_values."""
    arr = (C.c_uint32 * len(ids))(*ids)
    n, cub = C.c_size_t(), C.c_size_t()
    self.api.call("debug_codegen_scan", self._h, arr, len(ids), mode, 0, None, 0, C.byref(n), C.byref(cub))
    buf = C.create_string_buffer(n.value + 1)
    self.api.call("debug_codegen_scan", self._h, arr, len(ids), mode, 1 if compile else 0, buf, n.value + 1, C.byref(n), C.byref(cub))
    return buf.value.decode(), cub.value


Ir.debug_codegen_scan = _debug_codegen_scan


# convenience: named binary ops exactly as the reference's bop! expansion (internal.rs:218-227)
def _mk(kind):
    def f(self, lhs: int, rhs: int) -> int:
        return self.bop(kind, lhs, rhs)
    return f


for _n, _k in (("add", Bop.Add), ("sub", Bop.Sub), ("mul", Bop.Mul), ("div", Bop.Div), ("lt", Bop.Lt),
               ("gt", Bop.Gt), ("eq", Bop.Eq), ("leq", Bop.Leq), ("geq", Bop.Geq), ("neq", Bop.Neq),
               ("and_", Bop.And), ("or_", Bop.Or), ("xor", Bop.Xor), ("shl", Bop.Shl), ("shr", Bop.Shr),
               ("minimum", Bop.Min), ("maximum", Bop.Max)):
    setattr(Ir, _n, _mk(_k))


# -- backend lifecycle (product) ---------------------------------------------
def init(device: int = -1):
    product_api().call("init", device)


def shutdown():
    product_api().call("shutdown")


def sync():
    product_api().call("sync")


def stream_ptr() -> int:
    p = C.c_void_p()
    product_api().call("stream", C.byref(p))
    return p.value or 0


def device_index() -> int:
    d = C.c_int32()
    product_api().call("device", C.byref(d))
    return d.value


def stats() -> dict:
    s = Stats()
    product_api().call("stats", C.byref(s))
    return s.as_dict()


def stats_reset():
    product_api().call("stats_reset")


def cache_clear():
    product_api().call("cache_clear")
