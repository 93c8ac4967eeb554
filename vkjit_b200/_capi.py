"""ctypes binding of the C ABI declared in include/vkjit_b200.h.

`CApi(path, prefix)` binds one shared library.  The product only ever binds its
own `libvkjit_b200.so` (see `product_api()`); the parity tests bind the CPU
oracle with the same class so both sides are driven through identical host
code.  Nothing in this package references the oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
PRODUCT_LIB = os.environ.get("VKJIT_B200_LIB") or os.path.join(_HERE, "libvkjit_b200.so")   # override: A/B of builds

# status codes (include/vkjit_b200.h)
OK, ERR_INVALID, ERR_TYPE, ERR_SIZE, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA, ERR_COMPILE, ERR_DIST = range(9)


class VkjitError(RuntimeError):
    """The reference panics; the C ABI returns a status; Python raises."""

    def __init__(self, status: int, msg: str):
        super().__init__(f"[status {status}] {msg}")
        self.status = status


class VkjitTypeError(VkjitError, TypeError):
    pass


class VkjitSizeError(VkjitError, ValueError):
    pass


class VkjitNoDeviceError(VkjitError):
    pass


_u32, _i32, _sz, _u64, _f32 = C.c_uint32, C.c_int32, C.c_size_t, C.c_uint64, C.c_float
_p = C.c_void_p
_pu32, _psz, _pi32, _pu64 = C.POINTER(C.c_uint32), C.POINTER(C.c_size_t), C.POINTER(C.c_int32), C.POINTER(C.c_uint64)


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "cache_hits", "cache_misses", "trace_launches", "prim_launches", "last_compile_ns",
        "last_eval_ns", "bytes_h2d", "bytes_d2h", "pool_bytes_live", "collectives", "disk_hits")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


# name -> argtypes; every function returns int32 status unless listed in _SPECIAL
_SIGS = {
    # Ir-level API shared by the product and the oracle
    "ir_create": [C.POINTER(_p)],
    "ir_destroy": [_p],
    "type_struct": [_p, _pu32, _sz, _pu32],
    "type_struct_len": [_p, _u32, _psz],
    "type_struct_elem": [_p, _u32, _sz, _pu32],
    "const_f32": [_p, _f32, _pu32],
    "const_i32": [_p, _i32, _pu32],
    "const_u32": [_p, _u32, _pu32],
    "const_bool": [_p, _i32, _pu32],
    "array_f32": [_p, _p, _sz, _pu32],
    "array_i32": [_p, _p, _sz, _pu32],
    "array_u32": [_p, _p, _sz, _pu32],
    "array_bool": [_p, _p, _sz, _pu32],
    "array_empty": [_p, _u32, _sz, _pu32],
    "arange": [_p, _u32, _sz, _pu32],
    "linspace": [_p, _u32, _u32, _u32, _sz, _pu32],
    "zeros": [_p, _u32, _pu32],
    "ones": [_p, _u32, _pu32],
    "cast": [_p, _u32, _u32, _pu32],
    "bop": [_p, _i32, _u32, _u32, _pu32],
    "uop": [_p, _i32, _u32, _pu32],
    "bitcast": [_p, _u32, _u32, _pu32],
    "select": [_p, _u32, _u32, _u32, _pu32],
    "struct_init": [_p, _pu32, _sz, _pu32],
    "getattr": [_p, _u32, _sz, _pu32],
    "setattr": [_p, _u32, _u32, _sz, _pu32],
    "gather": [_p, _u32, _u32, _i32, _u32, _pu32],
    "scatter": [_p, _u32, _u32, _u32, _i32, _u32, _pu32],
    "scatter_add": [_p, _u32, _u32, _u32, _i32, _u32, _pu32],
    "var_type": [_p, _u32, _pu32],
    "var_ref_count": [_p, _u32, _pu32],
    "var_count": [_p, _psz],
    "var_deps": [_p, _u32, _pu32, _sz, _psz, _pi32, _pu32],
    "array_count": [_p, _psz],
    "is_buffer": [_p, _u32, _pi32],
    "var_size": [_p, _u32, _psz],
    "inc_ref": [_p, _u32],
    "dec_ref": [_p, _u32],
    "ir_repr": [_p, _p, _sz, _psz],
    "var_repr": [_p, _u32, _p, _sz, _psz],
    "schedule": [_p, _pu32, _sz],
    "eval": [_p, _pu32, _sz],
    "read": [_p, _u32, _u32, _p, _sz],
    "reduce": [_p, _i32, _u32, _pu32],
    "prefix_sum": [_p, _u32, _i32, _pu32],
    "compress": [_p, _u32, _pu32, _psz],
    "compress_values": [_p, _u32, _u32, _pu32, _psz],
    "shard_range": [_sz, _i32, _i32, _psz, _psz],
}
# product-only entry points
_PRODUCT_SIGS = {
    "init": [_i32],
    "shutdown": [],
    "stream": [C.POINTER(_p)],
    "device": [_pi32],
    "sync": [],
    "host_alloc": [_sz, C.POINTER(_p)],
    "host_free": [_p],
    "var_device_ptr": [_p, _u32, _pu64],
    "dist_unique_id": [_p],
    "dist_init": [_i32, _i32, _p],
    "dist_mailbox_handle": [_p],
    "dist_mailbox_open": [_p, _i32],
    "dist_set_p2p": [_i32],
    "dist_init_env": [],
    "debug_rendezvous": [_i32, _i32, C.c_char_p, _i32, _p, _p, _p, C.c_double],
    "dist_shutdown": [],
    "dist_info": [_pi32, _pi32],
    "arange_sharded": [_p, _u32, _sz, _pu32],
    "array_wrap_device": [_p, _u32, _u64, _sz, _pu32],
    "array_wrap_device_owned": [_p, _u32, _u64, _sz, _p, _p, _pu32],
    "var_to_dlpack": [_p, _u32, C.POINTER(_p)],
    "var_from_dlpack": [_p, _p, _pu32],
    "array_sharded": [_p, _u32, _p, _sz, _pu32],
    "array_shard_local": [_p, _u32, _p, _sz, _pu32],
    "var_is_sharded": [_p, _u32, _pi32],
    "var_shard_base": [_p, _u32, _pu64],
    "stats": [C.POINTER(Stats)],
    "stats_reset": [],
    "cache_clear": [],
    "debug_codegen": [_p, _pu32, _sz, _i32, _p, _sz, _psz, _psz],
    "debug_walk_ns": [_p, _pu32, _sz, _u32, _pu64, _pu32],
    "debug_eval_bookkeeping": [_p, _pu32, _sz],
    "debug_reduce_trace": [_pu64, _sz, _psz],
    "debug_codegen_reduce": [_p, _u32, _i32, _i32, _p, _sz, _psz, _psz],
    "debug_codegen_scan": [_p, _pu32, _sz, _i32, _i32, _p, _sz, _psz, _psz],
}
# oracle-only entry points (declared in oracle/oracle.h)
_ORACLE_SIGS = {
    "set_threads": [_i32],
    "var_host_ptr": [_p, _u32, C.POINTER(_p)],
    "arange_shard": [_p, _u32, _sz, _i32, _i32, _pu32],
    "fill_hash": [_p, _sz, _u64, _u32, _i32],
}


class CApi:
    def __init__(self, path: str, prefix: str):
        if not os.path.exists(path):
            raise ImportError(
                f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                "There is no fallback path.")
        self.path, self.prefix = path, prefix
        self.lib = C.CDLL(path, mode=C.RTLD_GLOBAL if prefix == "vkjit_" else C.RTLD_LOCAL)
        self._fn = {}
        for table in (_SIGS, _PRODUCT_SIGS, _ORACLE_SIGS):
            for name, argtypes in table.items():
                f = getattr(self.lib, prefix + name, None)
                if f is None:
                    continue
                f.argtypes, f.restype = argtypes, C.c_int32
                self._fn[name] = f
        le = getattr(self.lib, prefix + "last_error")
        le.argtypes, le.restype = [], C.c_char_p
        self._last_error = le

    def has(self, name: str) -> bool:
        return name in self._fn

    def last_error(self) -> str:
        s = self._last_error()
        return s.decode("utf-8", "replace") if s else ""

    def call(self, name: str, *args):
        st = self._fn[name](*args)
        if st != OK:
            msg = self.last_error()
            if st == ERR_TYPE:
                raise VkjitTypeError(st, msg)
            if st == ERR_SIZE:
                raise VkjitSizeError(st, msg)
            if st == ERR_NO_DEVICE:
                raise VkjitNoDeviceError(st, msg)
            raise VkjitError(st, msg)


_product = None


def product_api() -> CApi:
    """The one library the product ever loads."""
    global _product
    if _product is None:
        _product = CApi(PRODUCT_LIB, "vkjit_")
    return _product
