"""Synthetic inputs of SURVEY.md §8d, generated bit-identically on both sides: on the device as a
vkjit trace over `arange` (no host copies), in the oracle by orc_fill_hash straight into its array."""
import ctypes as C

from vkjit_b200.ir import Bop, VarType as T

RAW, UNIFORM, SIGNED_UNIFORM, IDX16, MASK = 0, 1, 2, 3, 4


def device_hash(ir, lanes, seed):
    c = ir.const_u32
    x = ir.bop(Bop.Xor, lanes, c(seed))
    s = ir.add(ir.mul(x, c(747796405)), c(2891336453))
    sh = ir.add(ir.bop(Bop.Shr, s, c(28)), c(4))
    w = ir.mul(ir.bop(Bop.Xor, ir.bop(Bop.Shr, s, sh), s), c(277803737))
    return ir.bop(Bop.Xor, ir.bop(Bop.Shr, w, c(22)), w)


def device_array(ir, n, seed, kind):
    """Evaluated device array of n lanes."""
    lanes = ir.arange(T.U32, n)
    h = device_hash(ir, lanes, seed)
    if kind == RAW:
        v = h
    elif kind == UNIFORM:
        v = ir.mul(ir.cast(ir.bop(Bop.Shr, h, ir.const_u32(8)), T.F32), ir.const_f32(2.0 ** -24))
    elif kind == SIGNED_UNIFORM:
        u = ir.mul(ir.cast(ir.bop(Bop.Shr, h, ir.const_u32(8)), T.F32), ir.const_f32(2.0 ** -24))
        v = ir.sub(ir.mul(u, ir.const_f32(2.0)), ir.const_f32(1.0))
    elif kind == IDX16:
        v = ir.bop(Bop.And, h, ir.const_u32(0xFFFF))
    else:
        v = ir.neq(ir.bop(Bop.And, h, ir.const_u32(1)), ir.const_u32(0))
    ir.eval([v])
    return v


def oracle_array(oir, n, seed, kind):
    ty = {RAW: T.U32, UNIFORM: T.F32, SIGNED_UNIFORM: T.F32, IDX16: T.U32, MASK: T.Bool}[kind]
    v = oir.array_empty(ty, n)
    p = C.c_void_p()
    oir.api.call("var_host_ptr", oir._h, v, C.byref(p))
    oir.api.call("fill_hash", p, n, 0, seed, kind)
    return v
