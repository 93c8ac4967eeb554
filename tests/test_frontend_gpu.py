"""The vkjit Python front-end mirror (`vkjit_b200.vkjit`: Var, eval, var, ir, linspace) on the CUDA path,
and the Monte-Carlo mega-trace (BASELINE configs[4]) against the oracle."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


@pytest.fixture()
def vkjit(cuda_backend):
    from vkjit_b200 import vkjit as m
    return m


def test_reference_python_surface(vkjit):
    """libs/vkjit-python/src/lib.rs:18-26 — Var, eval, var, ir, linspace; lazy __str__ (types.rs:148-154)."""
    x = vkjit.linspace(2.0, 4.0, 4)
    assert "Bop" in repr(x)                      # not evaluated yet: Debug of the node (types.rs:140-147)
    assert str(x) == "[2.0, 2.5, 3.0, 3.5]"      # __str__ evaluates on demand
    assert repr(x) == "array(dtype = F32, [2.0, 2.5, 3.0, 3.5])"
    assert x.tolist() == [2.0, 2.5, 3.0, 3.5]
    a = vkjit.Var([1.0, 2.0, 3.0])
    b = vkjit.var(1.0, 1.0, 1.0)                 # several args form one sequence (functions.rs:25-34)
    c = a + b
    d = (a - 1.0) * b / 2.0
    vkjit.eval([c, d])                           # eval([...]) works (the reference calls an unexported .id())
    assert c.tolist() == [2.0, 3.0, 4.0] and d.tolist() == [0.0, 0.5, 1.0]
    assert "Var {" in vkjit.ir()
    with pytest.raises(TypeError):
        vkjit.Var("nope")                        # "Not a valid argument!" (types.rs:81)
    with pytest.raises(AssertionError):
        vkjit.linspace(1, 2.0, 3)                # functions.rs:46 assert_eq!(start.ty(), stop.ty())


def test_coercion_order(vkjit):
    """types.rs:50-81: Var, u32, i32, f32, bool, [u32], [i32], [f32]"""
    from vkjit_b200 import VarType as T
    assert vkjit.var(3).ty() == T.U32 and vkjit.var(-3).ty() == T.I32 and vkjit.var(2.0).ty() == T.F32
    assert vkjit.var(True).ty() == T.U32         # a Python bool extracts as u32 first
    assert vkjit.var([1, 2]).ty() == T.U32 and vkjit.var([1, -2]).ty() == T.I32 and vkjit.var([1.0, 2]).ty() == T.F32
    z = vkjit.var([1, 2]) + (-1)                 # autocast U32 + I32 -> I32 (test.rs:176-187)
    assert z.ty() == T.I32 and z.tolist() == [0, 1]


def test_var_lifetime_releases_device_memory(vkjit, cuda_backend):
    """Clone/Drop own one reference each (types.rs:87-98): dropping the last handle frees the array."""
    import gc
    gc.collect()
    base = cuda_backend.stats()["pool_bytes_live"]
    x = vkjit.Var(np.ones(1 << 20, np.float32))
    y = x * 2.0
    vkjit.eval([y])
    assert cuda_backend.stats()["pool_bytes_live"] >= base + 2 * (4 << 20)
    del x, y
    gc.collect()
    assert cuda_backend.stats()["pool_bytes_live"] == base


def test_frontend_extensions(vkjit):
    from vkjit_b200 import VarType as T
    i = vkjit.arange(T.U32, 1000)
    m = (i & 1).eq(0)
    idx, cnt = m.compress()
    assert cnt == 500 and idx.tolist()[:3] == [0, 2, 4]
    assert i.sum().tolist() == [499500] and i.max().tolist() == [999]
    assert i.prefix_sum().tolist()[:4] == [0, 0, 1, 3]
    bins = vkjit.Var(np.zeros(10, np.uint32))
    one = vkjit.var(1)
    one.scatter_add(bins, i / 100)
    vkjit.eval([one])
    assert bins.tolist() == [100] * 10
    t = vkjit.select(i < 3, i.cast(T.F32), -1.0)
    assert t.tolist()[:5] == [0.0, 1.0, 2.0, -1.0, -1.0]


def test_monte_carlo_mega_trace_matches_oracle(vkjit, oir, cuda_backend):
    """~215-op trace, PCG + Box-Muller + exp + masked select.  Integer RNG state is bit-exact; the f32
    result accumulates 2-ulp transcendental differences -> relative tolerance 2e-5 (parity unpinned)."""
    import monte_carlo
    from ir_adapter import IrModule
    n, rounds = 1 << 14, 5
    y = monte_carlo.build(vkjit, n, rounds)
    got = y.numpy()
    om = IrModule(oir)
    yo = monte_carlo.build(om, n, rounds)
    oir.eval([yo.id])
    exp = oir.as_slice(yo.id, 5)
    assert got.shape == exp.shape and np.isfinite(got).all()
    assert np.allclose(got, exp, rtol=2e-5, atol=1e-5), float(np.abs(got - exp).max())
    # cached: building the same program again must not recompile
    cuda_backend.stats_reset()
    y2 = monte_carlo.build(vkjit, n, rounds)
    vkjit.eval([y2])
    st = cuda_backend.stats()
    assert st["cache_hits"] == 1 and st["cache_misses"] == 0


def test_zero_copy_interop_and_unaligned_views(vkjit, cuda_backend):
    """Foreign device memory is wrapped without a copy (CUDA Array Interface); a view that is only 4-byte
    aligned takes the scalar kernel variant and still matches; results export back to torch zero-copy."""
    import torch
    t = torch.arange(0, 4099, device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    x = vkjit.from_cuda_array(t)
    y = x * 2.0 + 1.0
    assert np.array_equal(y.numpy(), (t * 2 + 1).cpu().numpy())
    view = t[1:]                                  # base + 4 bytes: not 16-byte aligned
    assert view.data_ptr() % 16 != 0
    z = vkjit.from_cuda_array(view) * 0.5
    assert np.array_equal(z.numpy(), (view * 0.5).cpu().numpy())
    assert vkjit.from_cuda_array(view).sum().tolist() == [float(view.sum().item())]
    cuda_backend.sync()
    back = torch.as_tensor(y, device="cuda")      # zero-copy export
    assert back.data_ptr() == y.__cuda_array_interface__["data"][0]
    assert torch.equal(back, t * 2 + 1)
    # dropping the Var does not free the foreign memory
    del x, z
    import gc; gc.collect()
    assert float(t[5].item()) == 5.0


def test_dlpack_round_trip_with_torch(vkjit, cuda_backend):
    """DLPack in and out, zero-copy, with correct ownership in both directions."""
    import gc
    import torch
    # export: torch sees the Var's memory; the tensor keeps the array alive after the Var is gone
    y = vkjit.arange(vkjit.VarType.U32, 5000) * 3 + 1
    t = torch.from_dlpack(y)
    assert t.device.type == "cuda" and t.shape == (5000,) and t.data_ptr() == y.__cuda_array_interface__["data"][0]
    expect = np.arange(5000, dtype=np.uint32) * 3 + 1
    assert np.array_equal(t.cpu().numpy().view(np.uint32), expect)
    live = cuda_backend.stats()["pool_bytes_live"]
    del y; gc.collect()
    assert cuda_backend.stats()["pool_bytes_live"] == live            # still referenced by the exported tensor
    assert np.array_equal(t.cpu().numpy().view(np.uint32), expect)
    del t; gc.collect()
    assert cuda_backend.stats()["pool_bytes_live"] < live             # deleter ran: the var and its array are gone
    # re-evaluating an exported var replaces the var's array (eval of a Binding root copies, internal.rs:1192-1205);
    # the exported tensor keeps the OLD array: nothing may recycle it while torch still reads it
    y = vkjit.arange(vkjit.VarType.U32, 5000) * 3 + 1
    t = torch.from_dlpack(y)
    vkjit.eval([y])
    junk = [vkjit.arange(vkjit.VarType.U32, 5000) + 77 for _ in range(6)]   # same-size blocks: would reuse a recycled one
    vkjit.eval(junk)
    cuda_backend.sync()
    assert np.array_equal(t.cpu().numpy().view(np.uint32), expect)
    assert np.array_equal(y.numpy(), expect)
    del junk, y, t; gc.collect()
    # import: a Var over torch memory; the torch tensor object may die, the memory must not
    src = torch.arange(0, 4096, device="cuda", dtype=torch.float32)
    ptr = src.data_ptr()
    torch.cuda.synchronize()
    x = vkjit.from_dlpack(src)
    assert x.__cuda_array_interface__["data"][0] == ptr
    del src; gc.collect()
    junk = [torch.empty(4096, device="cuda") for _ in range(8)]      # would reuse the block if it had been freed
    assert all(j.data_ptr() != ptr for j in junk)
    assert np.array_equal((x + 1.0).numpy(), np.arange(4096, dtype=np.float32) + 1)
    del x; gc.collect()
    # unsupported tensors are rejected and stay usable by their owner
    with pytest.raises(Exception):
        vkjit.from_dlpack(torch.zeros(4, 4, device="cuda"))
    with pytest.raises(Exception):
        vkjit.from_dlpack(torch.zeros(8, device="cuda", dtype=torch.float64))
    with pytest.raises(Exception):
        vkjit.from_dlpack(torch.zeros(8, dtype=torch.float32))       # host memory


def test_rust_front_end_surface(vkjit):
    """The names vkjit-rust exposes (types.rs:142-197, functions.rs:5-82) exist with the same meaning; the two
    front-end tests of the reference run as written: `setattr` (types.rs:214-228) and `test_scatter` (:230-242)."""
    F32, U32 = vkjit.VarType.F32, vkjit.VarType.U32
    st = vkjit.zeros(vkjit._global_ir().struct_type([F32, F32]))
    x = vkjit.var([1.0, 2.0, 3.0])
    st.setattr(x, 0)
    a, b = st.getattr(0), st.getattr(1)
    vkjit.eval([a, b])
    assert a.to_vec() == [1.0, 2.0, 3.0] and b.to_vec() == [0.0, 0.0, 0.0]
    y = vkjit.var(7.0)
    target = vkjit.var([1.0, 2.0, 3.0])
    y.scatter(target, vkjit.arange(U32, 3))
    vkjit.eval([y])
    assert target.to_vec() == [7.0, 7.0, 7.0]
    # scatter_with / gather_with: masked forms
    src = vkjit.var([10, 20, 30, 40])
    idx = vkjit.arange(U32, 4)
    keep = idx.lt(2)
    g = vkjit.gather_with(src, idx, keep)
    dst = vkjit.var([0, 0, 0, 0])
    s = vkjit.var([5, 6, 7, 8])
    s.scatter_with(dst, idx, keep)
    vkjit.schedule([g])                       # schedule! then eval! of another var evaluates both
    vkjit.eval([s])
    assert g.to_vec() == [10, 20, 0, 0] and dst.to_vec() == [5, 6, 0, 0]
    p = vkjit.struct(src, vkjit.var(2.5))     # Var::from(&[Var])
    q = p.getattr(0) + vkjit.ones(U32)
    assert q.tolist() == [11, 21, 31, 41]
    assert vkjit.repr_ir() == vkjit.ir() and vkjit.repr_ir().startswith("Ir {")
