"""The native `vkjit` Python front-end (vkjit_b200/csrc/pyfront.cpp + vkjit_b200/vkjit.py) without a device: trace
construction, coercion order (vkjit-python/src/types.rs:47-82), operator lowering, ownership (types.rs:87-98) and
error mapping.  Building a trace needs no GPU; evaluating one does, and must fail loudly here."""
import gc
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from vkjit_b200 import VarType as T  # noqa: E402
from vkjit_b200 import vkjit  # noqa: E402
from vkjit_b200.ir import Ir  # noqa: E402


def live_vars():
    text = vkjit.ir()
    return text.count("Var {") - text.count("op: Free")


def test_native_module_is_what_runs():
    from vkjit_b200 import _native
    assert _native.__file__.endswith(".so")
    assert issubclass(vkjit.Var, _native.VarBase)
    x = vkjit.arange(T.F32, 4) * 2.0
    assert type(x) is vkjit.Var            # operators return the Python-visible class


def test_scalar_coercion_order():
    """Var, u32, i32, f32, bool (types.rs:50-66); a Python bool extracts as u32 first."""
    assert vkjit.var(3).ty() == T.U32 and vkjit.var(-3).ty() == T.I32 and vkjit.var(2.0).ty() == T.F32
    assert vkjit.var(True).ty() == T.U32 and vkjit.Var(False).ty() == T.U32
    assert vkjit.var(0xFFFFFFFF).ty() == T.U32 and vkjit.var(-2 ** 31).ty() == T.I32
    assert vkjit.var(np.float32(1.5)).ty() == T.F32 and vkjit.var(np.uint32(7)).ty() == T.U32
    assert vkjit.var(np.int64(-7)).ty() == T.I32 and vkjit.var(np.float64(0.1)).ty() == T.F32
    for bad in (2 ** 32, -2 ** 31 - 1, 10 ** 30, "nope", None, object(), 1j):
        with pytest.raises(TypeError, match="Not a valid argument!"):
            vkjit.Var(bad)
    c = vkjit.var(2.5)
    assert repr(c) == "Var { op: Const(Float32(2.5)), deps: [], side_effects: [], ty: F32, ref_count: 1 }"
    assert repr(vkjit.var(-3)) == "Var { op: Const(Int32(-3)), deps: [], side_effects: [], ty: I32, ref_count: 1 }"


def test_operators_lower_to_the_same_nodes_as_the_ir_api():
    x = vkjit.arange(T.F32, 8)
    i = vkjit.arange(T.U32, 8)

    def node(v):
        return repr(v).split(",")[0]

    assert node(x + 1.0) == "Var { op: Bop(Add)" and node(x - 1.0) == "Var { op: Bop(Sub)"
    assert node(x * x) == "Var { op: Bop(Mul)" and node(x / 2.0) == "Var { op: Bop(Div)"
    assert node(x.__div__(2.0)) == "Var { op: Bop(Div)"          # the reference's Python-2 name (types.rs:136-139)
    assert node(i & 1) == "Var { op: Bop(And)" and node(i | 1) == "Var { op: Bop(Or)" and node(i ^ 1) == "Var { op: Bop(Xor)"
    assert node(i << 1) == "Var { op: Bop(Shl)" and node(i >> 1) == "Var { op: Bop(Shr)"
    assert node(-x) == "Var { op: Uop(Neg)" and node(abs(x)) == "Var { op: Uop(Abs)" and node(~i) == "Var { op: Uop(Not)"
    for v, name in ((x < 1.0, "Lt"), (x > 1.0, "Gt"), (x <= 1.0, "Leq"), (x >= 1.0, "Geq"),
                    (x.lt(1.0), "Lt"), (x.gt(1.0), "Gt"), (x.eq(1.0), "Eq"), (x.leq(1.0), "Leq"), (x.geq(1.0), "Geq"),
                    (x.neq(1.0), "Neq")):
        assert node(v) == f"Var {{ op: Bop({name})" and v.ty() == T.Bool
    # reflected forms keep the operand order: const first
    y = 2.0 - x
    assert repr(y).startswith("Var { op: Bop(Sub), deps: [") and y.ty() == T.F32
    lhs, rhs = [int(t) for t in repr(y).split("deps: [")[1].split("]")[0].split(",")]
    assert rhs == x.id() and lhs != x.id()
    z = 1.0 < x                                                  # Python reflects `<` into x.__gt__(1.0)
    assert node(z) == "Var { op: Bop(Gt)"
    # autocast U32 + I32 -> I32 through an inserted Cast (internal.rs:146-166)
    assert (i + (-1)).ty() == T.I32 and (i + 1).ty() == T.U32 and (i + 1.0).ty() == T.F32
    # NumPy scalars on the left defer to Var (no element-wise broadcasting over an opaque object)
    w = np.float32(2.0) * x
    assert type(w) is vkjit.Var and node(w) == "Var { op: Bop(Mul)" and w.ty() == T.F32
    assert (np.uint32(3) + i).ty() == T.U32 and (np.int64(-3) + i).ty() == T.I32
    # `==` stays identity, so a Var is hashable and usable as a dict key
    assert (x == x) is True and (x == y) is False and len({x: 1, y: 2}) == 2
    assert node(vkjit.select(x < 1.0, x, 0.0)) == "Var { op: Select" and node(x.then_else(x, x)) == "Var { op: Select"
    assert node(vkjit.maximum(x, 0.0)) == "Var { op: Bop(Max)" and node(vkjit.minimum(0.0, x)) == "Var { op: Bop(Min)"
    assert node(vkjit.sqrt(x)) == "Var { op: Uop(Sqrt)" and node(vkjit.exp(2.0)) == "Var { op: Uop(Exp)"
    assert i.cast(T.F32).ty() == T.F32 and i.bitcast(T.F32).ty() == T.F32


def test_same_program_as_the_plain_ir_path():
    """The Monte-Carlo trace built through the native operators generates the same CUDA C as the same program
    replayed call by call through the `Ir` mirror (ctypes) — i.e. the same canonical trace."""
    import monte_carlo
    from ir_adapter import IrModule
    y = monte_carlo.build(vkjit, 1 << 12, 3)
    src_native, _ = vkjit._global_ir().debug_codegen([y.id()])
    plain = Ir()
    yp = monte_carlo.build(IrModule(plain), 1 << 12, 3)
    src_plain, _ = plain.debug_codegen([yp.id])
    plain.close()
    assert src_native == src_plain and "vkjit_trace" in src_native


def test_ownership_clone_drop_steal():
    gc.collect()
    base = live_vars()
    g = vkjit._global_ir()
    x = vkjit.arange(T.U32, 16)
    assert live_vars() == base + 1 and g.ref_count(x.id()) == 1
    y = x + 1                       # the const is owned by the Bop alone
    assert live_vars() == base + 3 and g.ref_count(x.id()) == 2
    c = x._clone()
    assert c.id() == x.id() and g.ref_count(x.id()) == 3
    v = vkjit.Var(x)                # Var(Var) clones (types.rs:50-52)
    assert v.id() == x.id() and g.ref_count(x.id()) == 4
    del c, v
    assert g.ref_count(x.id()) == 2
    same = x.cast(T.U32)            # Ir::cast returns the operand itself; the wrapper owns a new count
    assert same.id() == x.id() and g.ref_count(x.id()) == 3
    del same
    i = y._steal()                  # ownership moves out: the object no longer releases
    assert y.id() is None and y._id is None
    del y
    assert live_vars() == base + 3
    z = vkjit.Var._own(i)           # ... and back in
    assert z.id() == i
    with pytest.raises(TypeError):
        vkjit.Var.__new__(vkjit.Var) + 1   # a Var that owns nothing is not an operand
    del z, x
    gc.collect()
    assert live_vars() == base      # the whole trace was released


def test_errors_map_to_python_exceptions():
    from vkjit_b200._capi import VkjitError, VkjitNoDeviceError, VkjitTypeError
    x = vkjit.arange(T.F32, 8)
    with pytest.raises(VkjitTypeError) as e:          # select asserts lhs_ty == rhs_ty (internal.rs:232)
        vkjit.select(x < 1.0, x, 1)
    assert isinstance(e.value, TypeError) and e.value.status == 2
    with pytest.raises(TypeError, match="Not a valid argument!"):
        x + "a"
    with pytest.raises(TypeError):
        vkjit.eval([1, 2])
    with pytest.raises(VkjitError):
        vkjit.Var._own(0xFFFFFF).ty()
    import torch
    if not torch.cuda.is_available():                 # no CPU path: evaluating without a B200 fails loudly
        with pytest.raises(VkjitNoDeviceError):
            vkjit.eval([x * 2.0])
        with pytest.raises(VkjitNoDeviceError):
            vkjit.var([1.0, 2.0])                     # uploads need the device too


def test_trace_building_is_native_speed():
    """~0.2 us per node natively vs ~3 us through ctypes; the bound here is generous (shared CI cores)."""
    import time

    import monte_carlo
    best = 1e9
    for _ in range(20):
        t0 = time.perf_counter()
        y = monte_carlo.build(vkjit, 1 << 20, 5)
        best = min(best, time.perf_counter() - t0)
        del y
    assert best < 1.0e-3, best


def test_threads_can_build_traces_concurrently():
    import threading
    errs = []

    def work():
        try:
            for _ in range(200):
                a = vkjit.arange(T.U32, 64)
                b = ((a * 3 + 1) >> 1) ^ a
                assert b.ty() == T.U32
        except Exception as ex:  # pragma: no cover
            errs.append(ex)

    ts = [threading.Thread(target=work) for _ in range(4)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs


def test_closing_the_global_ir_detaches_live_handles():
    """Run last in this file: handles that outlive the global Ir must not touch it again; the next operation starts
    a fresh Ir."""
    x = vkjit.arange(T.U32, 8)
    y = x + 1
    stale = x * 3
    old = vkjit._global_ir()
    old.close()
    del y, x                                   # released into nothing: no crash, no error
    gc.collect()
    z = vkjit.arange(T.F32, 4) * 2.0           # lazily creates and binds a new global Ir
    assert vkjit._global_ir() is not old and z.ty() == T.F32
    assert repr(z).startswith("Var { op: Bop(Mul)")
    with pytest.raises(TypeError):
        stale + 1                              # a handle of the closed Ir is not an operand of the new one
    g = vkjit._global_ir()
    counts = [g.ref_count(z.id())]
    del stale                                  # ... and its release does not touch the new Ir's vars
    gc.collect()
    assert [g.ref_count(z.id())] == counts and "Var {" in vkjit.ir()
