"""Parity at BASELINE.json's full sizes (2^26 .. 2^28 lanes).  Integer / index / mask / compress work is
compared bit for bit against the multi-threaded CPU oracle on the whole array (inputs come from the
stateless hash generator on both sides, so nothing but results crosses PCIe); f32 sums use the stated
tolerance; size-independent properties (checksums, round trips) are asserted on top."""
import math
import os

import numpy as np
import pytest

from synth import IDX16, MASK, RAW, SIGNED_UNIFORM, UNIFORM, device_array, oracle_array
from trace_gen import same_bits
from vkjit_b200.ir import Bop, Red, VarType as T

pytestmark = pytest.mark.gpu
N28, N26 = 1 << 28, 1 << 26


@pytest.fixture()
def oir_mt(oir):
    oir.api.call("set_threads", os.cpu_count() or 1)
    yield oir
    oir.api.call("set_threads", 1)


def test_generator_is_bit_identical_on_both_sides(cir, oir_mt):
    n = (1 << 22) + 5
    for kind, ty in ((RAW, T.U32), (UNIFORM, T.F32), (SIGNED_UNIFORM, T.F32), (IDX16, T.U32), (MASK, T.Bool)):
        a = cir.as_slice(device_array(cir, n, 1234 + kind, kind), ty)
        b = oir_mt.as_slice(oracle_array(oir_mt, n, 1234 + kind, kind), ty)
        assert same_bits(a, b, ty == T.F32), kind


def test_R28_reductions_full_size(cir, oir_mt):
    """configs[1]: sum/min/max over 2^28 lanes — u32 bit-exact, f32 within 1e-6*log2(n) (sum) / exact (min, max)."""
    xu_d, xu_o = device_array(cir, N28, 0xB2000021, RAW), oracle_array(oir_mt, N28, 0xB2000021, RAW)
    for r in (Red.Sum, Red.Min, Red.Max):
        assert cir.as_slice(cir.reduce(r, xu_d), T.U32)[0] == oir_mt.as_slice(oir_mt.reduce(r, xu_o), T.U32)[0]
    cir.dec_ref_count(xu_d); oir_mt.dec_ref_count(xu_o)
    for kind in (UNIFORM, SIGNED_UNIFORM):
        xd, xo = device_array(cir, N28, 0xB2000022, kind), oracle_array(oir_mt, N28, 0xB2000022, kind)
        gs, os_ = (float(i.as_slice(i.reduce(Red.Sum, v), T.F32)[0]) for i, v in ((cir, xd), (oir_mt, xo)))
        scale = N28 * (0.5 if kind == UNIFORM else 0.5)          # sum of magnitudes
        assert abs(gs - os_) <= 1e-6 * 28 * scale, (gs, os_)
        for r in (Red.Min, Red.Max):
            assert cir.as_slice(cir.reduce(r, xd), T.F32)[0] == oir_mt.as_slice(oir_mt.reduce(r, xo), T.F32)[0]
        cir.dec_ref_count(xd); oir_mt.dec_ref_count(xo)


def test_E28_fused_elementwise_full_size(cir, oir_mt):
    """z = x*y + c over 2^28 f32: whole array bit-exact against the oracle (two roundings, no FMA)."""
    out = []
    for ir, mk in ((cir, device_array), (oir_mt, oracle_array)):
        x, y = mk(ir, N28, 0xB2000011, UNIFORM), mk(ir, N28, 0xB2000012, UNIFORM)
        z = ir.add(ir.mul(x, y), ir.const_f32(0.5))
        ir.eval([z])
        ir.dec_ref_count(x); ir.dec_ref_count(y)
        out.append(ir.as_slice(z, T.F32))
        ir.dec_ref_count(z)
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))


def test_C28_prefix_sum_and_compress_full_size(cir, oir_mt):
    """configs[3]: exclusive prefix sum and stream compaction of 2^28 u32 — bit-exact on the whole array,
    plus the checksum properties last + in[last] == sum and count == sum(mask)."""
    vd, vo = device_array(cir, N28, 0xB2000041, RAW), oracle_array(oir_mt, N28, 0xB2000041, RAW)
    sd, so = cir.prefix_sum(vd, True), oir_mt.prefix_sum(vo, True)
    a, b = cir.as_slice(sd, T.U32), oir_mt.as_slice(so, T.U32)
    assert np.array_equal(a, b)
    total = int(cir.as_slice(cir.reduce(Red.Sum, vd), T.U32)[0])
    # checksum property: exclusive[last] + in[last] == sum (mod 2^32)
    last = cir.add(cir.arange(T.U32, 1), cir.const_u32(N28 - 1))
    last_in = int(cir.as_slice_eval(cir.gather(vd, last), T.U32)[0])
    assert a[0] == 0 and (int(a[-1]) + last_in) % (1 << 32) == total
    del a, b
    cir.dec_ref_count(sd); oir_mt.dec_ref_count(so)
    md, mo = device_array(cir, N28, 0xB2000042, MASK), oracle_array(oir_mt, N28, 0xB2000042, MASK)
    (cd, nd), (co, no) = cir.compress_values(vd, md), oir_mt.compress_values(vo, mo)
    assert nd == no and abs(nd - N28 // 2) < (1 << 16)
    assert np.array_equal(cir.as_slice(cd, T.U32), oir_mt.as_slice(co, T.U32))
    # count == sum(mask) computed by an independent path (cast + fused trace-reduce)
    assert nd == int(cir.as_slice(cir.reduce(Red.Sum, cir.cast(md, T.U32)), T.U32)[0])
    cir.dec_ref_count(cd); oir_mt.dec_ref_count(co)
    (id_, n2), (io, n3) = cir.compress(md), oir_mt.compress(mo)
    assert n2 == nd == n3 and np.array_equal(cir.as_slice(id_, T.U32), oir_mt.as_slice(io, T.U32))
    # round trip: gathering the values at the compressed indices reproduces compress_values
    back = cir.gather(vd, id_)
    cir.eval([back])
    cd2, _ = cir.compress_values(vd, md)
    assert cir.size(id_) == nd and np.array_equal(cir.as_slice(back, T.U32), cir.as_slice(cd2, T.U32))


def test_C28_fused_mask_compress_full_size(cuda_backend, cir, oir_mt):
    """configs[3], fused-mask variant: the mask is a trace (`v > 2^31`, and a bit of a hash of v)
    evaluated inside the compaction kernel — bit-exact on the whole result; ONE kernel, nothing materialised."""
    vd, vo = device_array(cir, N28, 0xB2000041, RAW), oracle_array(oir_mt, N28, 0xB2000041, RAW)
    res = []
    for ir, v in ((cir, vd), (oir_mt, vo)):
        m1 = ir.gt(v, ir.const_u32(1 << 31))
        m2 = ir.neq(ir.bop(Bop.And, ir.mul(v, ir.const_u32(2654435761)), ir.const_u32(1 << 19)), ir.const_u32(0))
        if ir is cir:
            cuda_backend.stats_reset()
        (a, na), (b, nb) = ir.compress_values(v, m1), ir.compress(m2)
        if ir is cir:
            assert cuda_backend.stats()["trace_launches"] == 2 and not ir.is_buffer(m1) and not ir.is_buffer(m2)
        res.append((na, nb, ir.as_slice(a, T.U32), ir.as_slice(b, T.U32)))
        ir.dec_ref_count(a); ir.dec_ref_count(b)
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1]
    assert np.array_equal(res[0][2], res[1][2]) and np.array_equal(res[0][3], res[1][3])
    assert (res[0][2] > (1 << 31)).all() and (np.diff(res[0][3].astype(np.int64)) > 0).all()   # filter holds; indices ascend


def test_H26_skewed_indices_full_size(cir, oir_mt):
    """configs[2], skewed run: indices min(h & 0xFFFF, h >> 16) — low bins are hit more often."""
    out = []
    for ir, mk in ((cir, device_array), (oir_mt, oracle_array)):
        h = mk(ir, N26, 0xB2000031, RAW)
        idx = ir.bop(Bop.Min, ir.bop(Bop.And, h, ir.const_u32(0xFFFF)), ir.shr(h, ir.const_u32(16)))
        ir.eval([idx])
        table = mk(ir, 1 << 16, 0xB2000032, RAW)
        bins = ir.array_u32(np.zeros(1 << 16, np.uint32))
        s = ir.scatter_add(ir.gather(table, idx), bins, idx)
        ir.eval([s])
        out.append(ir.as_slice(bins, T.U32))
    assert np.array_equal(out[0], out[1])


def test_H26_gather_scatter_add_full_size(cir, oir_mt):
    """configs[2]: 2^26 indices into 2^16 bins, weighted (gather) and counting histograms — bit-exact;
    checksum: the bins of the counting histogram sum to the number of indices."""
    out = []
    for ir, mk in ((cir, device_array), (oir_mt, oracle_array)):
        idx = mk(ir, N26, 0xB2000031, IDX16)
        table = mk(ir, 1 << 16, 0xB2000032, RAW)
        b1 = ir.array_u32(np.zeros(1 << 16, np.uint32))
        b2 = ir.array_u32(np.zeros(1 << 16, np.uint32))
        s1 = ir.scatter_add(ir.gather(table, idx), b1, idx)
        s2 = ir.scatter_add(ir.const_u32(1), b2, idx)
        ir.eval([s1, s2])
        out.append((ir.as_slice(b1, T.U32), ir.as_slice(b2, T.U32)))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert int(out[0][1].astype(np.uint64).sum()) == N26


def test_M26_monte_carlo_full_size(cuda_backend, oir_mt):
    """configs[4]: the ~200-op Monte-Carlo mega-trace (364 IR nodes: PCG RNG, log, sqrt, sin/cos, exp, masked select)
    over 2^26 lanes — ALL 2^26 device results against the multi-threaded oracle, BIT-EXACT: integer ops are exact on
    both sides, f32 + - * / and sqrt are IEEE on both sides (no contraction), and exp/log/sin/cos are the shared
    vk_math.h.  No tolerance."""
    import monte_carlo
    from ir_adapter import IrModule
    from vkjit_b200 import vkjit
    y = monte_carlo.build(vkjit, N26, 5)
    got = y.numpy()
    assert got.shape == (N26,) and np.isfinite(got).all() and (got >= 0).all()
    yo = monte_carlo.build(IrModule(oir_mt), N26, 5)
    oir_mt.eval([yo.id])
    exp = oir_mt.as_slice(yo.id, T.F32)
    bad = np.nonzero(got.view(np.uint32) != exp.view(np.uint32))[0]
    assert bad.size == 0, (bad.size, bad[:4], got[bad[:4]], exp[bad[:4]])


def test_lane_indices_beyond_2_to_31(cuda_backend, cir):
    """Maximum sizes: n = 2^31 + 5 lanes (the reference's invocation index is 32-bit; so is ours, and nothing may
    wrap at 2^31).  No oracle here (8 GiB per array on the host): every result has a closed form."""
    n = (1 << 31) + 5
    M = 1 << 32
    i = cir.arange(T.U32, n)
    x = cir.add(cir.mul(i, cir.const_u32(3)), cir.const_u32(1))                 # x_i = 3 i + 1 (mod 2^32)
    want_sum = (3 * (n * (n - 1) // 2) + n) % M
    assert int(cir.as_slice(cir.reduce(Red.Sum, x), T.U32)[0]) == want_sum     # fused trace -> reduce, nothing stored
    assert int(cir.as_slice(cir.reduce(Red.Max, i), T.U32)[0]) == n - 1
    probe = np.array([0, 1, (1 << 31) - 1, 1 << 31, (1 << 31) + 1, n - 1], dtype=np.uint32)
    pv = cir.array_u32(probe)
    cir.eval([x])                                                              # 8 GiB elementwise store (vector + scalar tail)
    got = cir.as_slice_eval(cir.gather(x, pv), T.U32)
    assert np.array_equal(got, ((3 * probe.astype(np.uint64) + 1) % M).astype(np.uint32))
    assert int(cir.as_slice(cir.reduce(Red.Sum, x), T.U32)[0]) == want_sum     # hand-written reduce over 8 GiB
    s = cir.prefix_sum(x, True)                                                # look-back scan over 87382 tiles
    got = cir.as_slice_eval(cir.gather(s, pv), T.U32)
    p64 = probe.astype(object)
    assert [int(v) for v in got] == [int((3 * (p * (p - 1) // 2) + p) % M) for p in p64]
    cir.dec_ref_count(s)
    m = cir.eq(cir.bop(Bop.And, i, cir.const_u32(1023)), cir.const_u32(3))     # every 1024th lane, computed in-kernel
    idx, cnt = cir.compress(m)
    assert cnt == (n - 4) // 1024 + 1 and cir.size(idx) == cnt
    tail = cir.as_slice_eval(cir.gather(idx, cir.array_u32(np.array([0, cnt // 2, cnt - 1], dtype=np.uint32))), T.U32)
    assert [int(v) for v in tail] == [3, 3 + 1024 * (cnt // 2), 3 + 1024 * (cnt - 1)] and int(tail[-1]) == (1 << 31) + 3
