"""The reference's own known-answer tests, restated against the `Ir` mirror.

Each function follows one reference test line by line (file:line in the
docstring) and takes a fresh `Ir`; the same bodies run against the CPU oracle
(tests/test_oracle_golden.py, pins the oracle) and against the CUDA path
(tests/test_cuda_golden.py, `-m gpu`).  All comparisons are bit-exact.
"""
import numpy as np

from vkjit_b200.ir import VarType

F32, U32, I32, BOOL = VarType.F32, VarType.U32, VarType.I32, VarType.Bool


def _eq(a, expected, dtype):
    e = np.asarray(expected, dtype=dtype)
    assert a.dtype == e.dtype and a.shape == e.shape, (a, e)
    assert a.tobytes() == e.tobytes(), (a, e)


def test_linspace_f32(ir):
    """libs/vkjit-core/src/test.rs:10-20"""
    start = ir.const_f32(2.)
    stop = ir.const_f32(4.)
    x = ir.linspace(F32, start, stop, 4)
    ir.eval([x])
    _eq(ir.as_slice(x, F32), [2., 2.5, 3., 3.5], np.float32)


def test_linspace_eval2(ir):
    """libs/vkjit-core/src/test.rs:22-44 — two evals on one Ir; 3/10*10 is an exact RNE tie"""
    start = ir.const_f32(2.)
    stop = ir.const_f32(4.)
    x = ir.linspace(F32, start, stop, 4)
    ir.eval([x])
    _eq(ir.as_slice(x, F32), [2., 2.5, 3., 3.5], np.float32)
    start = ir.const_f32(10.)
    stop = ir.const_f32(20.)
    x = ir.linspace(F32, start, stop, 10)
    ir.eval([x])
    _eq(ir.as_slice(x, F32), [10., 11., 12., 13., 14., 15., 16., 17., 18., 19.], np.float32)


def test_add_f32(ir):
    """test.rs:47-58"""
    x = ir.arange(F32, 3)
    y = ir.arange(F32, 3)
    z = ir.add(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, F32), [0., 2., 4.], np.float32)


def test_add_u32(ir):
    """test.rs:60-71"""
    x = ir.arange(U32, 3)
    y = ir.arange(U32, 3)
    z = ir.add(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, U32), [0, 2, 4], np.uint32)


def test_add_i32(ir):
    """test.rs:73-84"""
    x = ir.arange(I32, 3)
    y = ir.arange(I32, 3)
    z = ir.add(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, I32), [0, 2, 4], np.int32)


def test_sub_f32(ir):
    """test.rs:87-98"""
    x = ir.array_f32([1., 2., 3.])
    y = ir.array_f32([0., 1., 2.])
    z = ir.sub(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, F32), [1., 1., 1.], np.float32)


def test_sub_u32(ir):
    """test.rs:100-111"""
    x = ir.array_u32([1, 2, 3])
    y = ir.array_u32([0, 1, 2])
    z = ir.sub(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, U32), [1, 1, 1], np.uint32)


def test_sub_i32(ir):
    """test.rs:113-124"""
    x = ir.array_i32([0, 1, 2])
    y = ir.array_i32([1, 2, 3])
    z = ir.sub(x, y)
    ir.eval([z])
    _eq(ir.as_slice(z, I32), [-1, -1, -1], np.int32)


def test_scatter_f32(ir):
    """test.rs:127-142 — read the scatter TARGET"""
    idx = ir.arange(U32, 3)
    x = ir.array_f32([0., 1., 2.])
    c = ir.const_f32(1.)
    x = ir.add(x, c)
    y = ir.array_f32([0., 0., 0.])
    x = ir.scatter(x, y, idx, None)
    ir.eval([x])
    _eq(ir.as_slice(y, F32), [1., 2., 3.], np.float32)
    # the scatter var itself is a root too and owns an output holding `src` (internal.rs:1076, :1192-1205)
    _eq(ir.as_slice(x, F32), [1., 2., 3.], np.float32)


def test_scatter_conditional(ir):
    """test.rs:144-161"""
    idx = ir.arange(U32, 5)
    x = ir.array_f32([0., 1., 2., 3., 4.])
    y = ir.array_f32([0., 0., 0., 0., 0.])
    const3 = ir.const_u32(3)
    cond = ir.lt(idx, const3)
    x = ir.scatter(x, y, idx, cond)
    ir.eval([x])
    _eq(ir.as_slice(y, F32), [0., 1., 2., 0., 0.], np.float32)


def test_cast_u32_to_f32(ir):
    """test.rs:164-173"""
    x = ir.arange(U32, 3)
    y = ir.cast(x, F32)
    ir.eval([y])
    _eq(ir.as_slice(y, F32), [0., 1., 2.], np.float32)


def test_autocast(ir):
    """test.rs:176-187 — U32 < I32 in the promotion order; U32->I32 is a bit reinterpret"""
    x = ir.array_u32([1, 2])
    y = ir.const_i32(-1)
    z = ir.add(x, y)
    assert ir.ty(z) == I32
    ir.eval([z])
    _eq(ir.as_slice(z, I32), [0, 1], np.int32)


def test_dec_ref_count(ir):
    """test.rs:190-207 — white-box: vars.len(), vars[0].ref_count, arrays.len()"""
    x = ir.array_f32([1., 2., 3.])
    c = ir.const_f32(1.)
    y = ir.add(x, c)
    ir.dec_ref_count(c)
    ir.dec_ref_count(x)
    ir.eval([y])
    assert ir.num_vars() == 3
    assert ir.ref_count(0) == 0
    assert ir.num_arrays() == 1
    _eq(ir.as_slice(y, F32), [2., 3., 4.], np.float32)


def test_setattr(ir):
    """libs/vkjit-rust/src/types.rs:214-228, through the Ir calls the Rust front-end makes:
    Var::from(vec) -> array_f32; zeros(Struct) -> Ir::zeros; setattr -> Ir::setattr then the old
    handle is dropped; getattr; eval!(st_1, st_2)."""
    x = ir.array_f32([1., 2., 3.])
    st_ty = ir.struct_type([F32, F32])
    st = ir.zeros(st_ty)
    ir.inc_ref_count(x)            # x.clone()
    st_new = ir.setattr(st, x, 0)  # Var::setattr (types.rs:152-159)
    ir.dec_ref_count(st)           # *self = ret drops the old struct handle
    ir.dec_ref_count(x)            # the temporary clone is dropped at the end of setattr
    st_1 = ir.getattr(st_new, 0)
    st_2 = ir.getattr(st_new, 1)
    ir.eval([st_1, st_2])
    _eq(ir.as_slice(st_1, F32), [1., 2., 3.], np.float32)
    _eq(ir.as_slice(st_2, F32), [0., 0., 0.], np.float32)


def test_scatter_const(ir):
    """libs/vkjit-rust/src/types.rs:230-242 — scatter a constant: y = 7.0; y.scatter(x, arange)"""
    x = ir.array_f32([1., 2., 3.])
    y = ir.const_f32(7.)
    idx = ir.arange(U32, 3)
    ir.inc_ref_count(x)  # x.clone()
    y2 = ir.scatter(y, x, idx, None)
    ir.dec_ref_count(y)
    ir.eval([y2])
    _eq(ir.as_slice(x, F32), [7., 7., 7.], np.float32)


def test_main_rs(ir):
    """src/main.rs:4-12 — arange(U32, 10); eval!(x); dbg!(x)"""
    x = ir.arange(U32, 10)
    ir.eval([x])
    _eq(ir.as_slice(x, U32), list(range(10)), np.uint32)
    assert ir.str(x) == "[0, 1, 2, 3, 4, 5, 6, 7, 8, 9]"


ALL = [test_linspace_f32, test_linspace_eval2, test_add_f32, test_add_u32, test_add_i32, test_sub_f32,
       test_sub_u32, test_sub_i32, test_scatter_f32, test_scatter_conditional, test_cast_u32_to_f32,
       test_autocast, test_dec_ref_count, test_setattr, test_scatter_const, test_main_rs]
