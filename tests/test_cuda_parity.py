"""Parity of the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): bit-exact for integer / index / mask / scatter / compress work and,
in practice, for f32 + - * / sqrt and casts (IEEE on both sides, no FMA contraction); f32 sums
within 1e-6*log2(n) relative of the f64-accumulated oracle; exp/log/sin/cos BIT-EXACT against the oracle (both
sides compile vkjit_b200/csrc/vk_math.h; the header itself is within 1 ulp of the exact value, tests/test_vk_math_cpu.py).
Everything outside the reference's own golden tests is "parity unpinned": the oracle is the spec.
"""
import math
import os

import numpy as np
import pytest

from trace_gen import BOOL, F32, I32, U32, TraceBuilder, same_bits, special_f32, special_u32, ulp_diff
from vkjit_b200.ir import Bop, Red, Uop

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def both(cir, oir):
    return (cir, oir)


def read(ir, v):
    return ir.as_slice(v, ir.ty(v))


# ---------------------------------------------------------------- fused elementwise traces
@pytest.mark.parametrize("seed", range(24))
@pytest.mark.parametrize("n", [1, 3, 4, 5, 1023, 4099])
def test_random_trace_bit_exact(cir, oir, seed, n):
    if n > 5 and seed >= 8:
        pytest.skip("large sizes: 8 seeds are enough")
    outs = []
    for ir in (cir, oir):
        tb = TraceBuilder(ir, seed * 7919 + n, n)
        roots = tb.build()
        ir.eval(roots)
        outs.append([(ir.ty(r), read(ir, r)) for r in roots])
    for (ty_c, a), (ty_o, b) in zip(*outs):
        assert ty_c == ty_o
        assert same_bits(a, b, ty_c == F32), (seed, n, ty_c, a[:8], b[:8])


def transcendental_inputs(seed, n=4096):
    """Every regime of vk_math.h: ordinary values, |x| up to 2^20 and far beyond (Payne-Hanek), values next to multiples
    of pi/2, subnormals, +-0, +-inf, NaN, the overflow / underflow thresholds of exp, arguments of log next to 1."""
    rng = np.random.default_rng(seed)
    parts = [
        rng.uniform(-20, 20, n // 8), rng.uniform(-105615, 105615, n // 8), rng.uniform(-2.0 ** 20, 2.0 ** 20, n // 8),
        rng.uniform(-1, 1, n // 8) * 10.0 ** rng.uniform(-45, 38, n // 8),           # all magnitudes incl. subnormals and huge
        (np.arange(n // 8) - n // 16) * (np.pi / 2) + rng.uniform(-1e-4, 1e-4, n // 8),  # near multiples of pi/2
        rng.uniform(80, 90, n // 16), rng.uniform(-110, -80, n // 16),                # exp: overflow / subnormal results
        1.0 + rng.uniform(-0.3, 0.45, n // 8),                                         # log: f = m - 1 over its whole range
    ]
    x = np.concatenate(parts).astype(np.float32)
    bits = rng.integers(0, 2 ** 32, n - len(x) - 12, dtype=np.uint64).astype(np.uint32).view(np.float32)  # arbitrary bit patterns (NaNs too)
    special = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, 88.72284, 88.72283, -103.97208, -103.9721, 1e-45, 105615.0], np.float32)
    return np.concatenate([x, bits, special])


@pytest.mark.parametrize("seed", range(6))
def test_transcendentals_bit_exact(cir, oir, seed):
    """exp / log / sin / cos (no reference op; the oracle is the specification): the generated kernels and the oracle
    compile the SAME vk_math.h, so the device must return the oracle's bits for EVERY input — 0 ulp, NaN patterns
    included; sqrt is IEEE on both sides.  (Accuracy of vk_math.h itself against f64 libm: <= 1 ulp, CPU tier.)"""
    xs = transcendental_inputs(seed)
    outs = []
    for ir in (cir, oir):
        x = ir.array_f32(xs)
        r = [ir.exp(x), ir.log(x), ir.sin(x), ir.cos(x), ir.sqrt(x)]
        ir.eval(r)
        outs.append([read(ir, v) for v in r])
        # sin and cos of the same operand in separate kernels: the unpaired vk_sinf / vk_cosf lowering
        s1, c1 = ir.sin(x), ir.cos(ir.add(x, ir.const_f32(0.0)))
        ir.eval([s1]); ir.eval([c1])
        outs[-1] += [read(ir, s1), read(ir, c1)]
    for name, a, b in zip(("exp", "log", "sin", "cos", "sqrt", "sin alone", "cos alone"), *outs):
        if name == "sqrt":
            assert same_bits(a, b, True), name
        else:
            bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
            assert bad.size == 0, (name, xs[bad[:4]], a[bad[:4]], b[bad[:4]])
    # x + 0.0 only differs from x for -0.0 -> +0.0 (cos is even): the two cos columns agree as well
    assert np.array_equal(outs[0][3].view(np.uint32), outs[0][6].view(np.uint32))


def test_range_proven_fast_paths_bit_exact(cir, oir):
    """Traces whose transcendental arguments have PROVABLE ranges take the unchecked vk_math.h fast paths in the generated
    kernel (program.cpp: FRange; which calls are proven is checked on the CPU tier) — the oracle always runs the checked
    functions.  Bit-exact over 2^20 pseudo-random arguments plus both endpoints of every range."""
    n = 1 << 20
    outs = []
    for ir in (cir, oir):
        lane = ir.arange(U32, n)
        u24 = ir.bop(Bop.Shr, ir.mul(lane, ir.const_u32(2654435761)), ir.const_u32(8))
        u24 = ir.select(ir.lt(lane, ir.const_u32(1)), ir.const_u32(0),
                        ir.select(ir.lt(lane, ir.const_u32(2)), ir.const_u32(0xFFFFFF), u24))   # lanes 0 / 1: the endpoints
        unit = ir.mul(ir.cast(u24, F32), ir.const_f32(2.0 ** -24))                               # [0, 1 - 2^-24]
        unit1 = ir.mul(ir.cast(ir.add(u24, ir.const_u32(1)), F32), ir.const_f32(2.0 ** -24))     # [2^-24, 1]
        th = ir.mul(unit, ir.const_f32(105615.0))
        rad = ir.sqrt(ir.mul(ir.log(unit1), ir.const_f32(-2.0)))
        r = [ir.log(unit1), ir.log(ir.mul(unit1, ir.const_f32(1e30))), ir.log(ir.mul(unit1, ir.const_f32(2.0 ** -102))),
             ir.exp(ir.sub(ir.mul(unit, ir.const_f32(173.8)), ir.const_f32(86.9))),
             ir.sin(th), ir.cos(th), ir.sin(ir.mul(unit, ir.const_f32(6.2831854820251465))),
             ir.exp(ir.add(ir.mul(ir.mul(rad, ir.cos(th)), ir.const_f32(0.2)), ir.const_f32(0.01))),
             ir.exp(ir.exp(ir.mul(unit, ir.const_f32(4.4))))]
        ir.eval(r)
        outs.append([read(ir, v) for v in r])
    for k, (a, b) in enumerate(zip(*outs)):
        bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
        assert bad.size == 0, (k, bad[:4], a[bad[:4]], b[bad[:4]])


@pytest.mark.parametrize("block", range(4))
def test_ranged_random_traces_bit_exact(cir, oir, block):
    """Random traces over lane-index leaves and the ops the range analysis models (the generator of the CPU-tier
    soundness test): about half of their exp / log / sin / cos calls are range-proven fast paths in the generated kernel.
    Every root bit-exact against the oracle (which always runs the checked functions)."""
    from test_product_cpu import _ranged_trace
    n = 4099
    for seed in range(block * 12, block * 12 + 12):
        outs = []
        for ir in (cir, oir):
            roots = _ranged_trace(ir, seed, n)
            ir.eval(roots)
            outs.append([read(ir, r) for r in roots])
            for r in roots:
                ir.dec_ref_count(r)
        for k, (a, b) in enumerate(zip(*outs)):
            bad = np.nonzero(a.view(np.uint32) != b.view(np.uint32))[0]
            # NaN payloads may differ between the host's and the device's f32 arithmetic (not in vk_math.h, which returns one pattern)
            bad = [i for i in bad if not (a.dtype == np.float32 and np.isnan(a[i]) and np.isnan(b[i]))]
            assert len(bad) == 0, (seed, k, bad[:4], a[bad[:4]], b[bad[:4]])


def test_large_elementwise_and_cache_hit(cuda_backend, cir, oir):
    """x*y+c at 2^20 (config-1 shape); second eval of the same structure must be a cache hit."""
    n = 1 << 20
    rng = np.random.default_rng(1)
    xs, ys = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    res = []
    for ir in (cir, oir):
        x, y, c = ir.array_f32(xs), ir.array_f32(ys), ir.const_f32(0.5)
        z = ir.add(ir.mul(x, y), c)
        ir.eval([z])
        res.append(read(ir, z))
    assert same_bits(res[0], res[1], True)
    # the reference semantics are two roundings (OpFMul then OpFAdd), not an FMA
    assert same_bits(res[0], (xs * ys + np.float32(0.5)).astype(np.float32), True)
    cuda_backend.stats_reset()
    x, y, c = cir.array_f32(ys), cir.array_f32(xs), cir.const_f32(0.5)
    z = cir.add(cir.mul(x, y), c)
    cir.eval([z])
    st = cuda_backend.stats()
    assert st["cache_hits"] == 1 and st["cache_misses"] == 0, st


def test_ragged_sizes_and_tail(cir, oir):
    for n in (1, 2, 3, 5, 6, 7, 255, 257, 1025):
        res = []
        for ir in (cir, oir):
            a = ir.array_u32(np.arange(n, dtype=np.uint32) * 3)
            z = ir.add(ir.mul(a, ir.arange(U32, n)), ir.const_u32(7))
            ir.eval([z])
            res.append(read(ir, z))
        assert same_bits(res[0], res[1], False), n


# ---------------------------------------------------------------- gather / scatter / scatter_add
@pytest.mark.parametrize("n,m", [(7, 3), (1000, 64), (65536, 1024), (100003, 65536)])
def test_gather_masked_and_unmasked(cir, oir, n, m):
    rng = np.random.default_rng(n)
    table = special_u32(rng, m)
    idx = rng.integers(0, m, n).astype(np.uint32)
    res = []
    for ir in (cir, oir):
        t, i = ir.array_u32(table), ir.array_u32(idx)
        g = ir.gather(t, i)
        act = ir.lt(i, ir.const_u32(m // 2))
        gm = ir.gather(t, i, act)
        tf = ir.array_f32(table.view(np.float32))
        gf = ir.gather(tf, i, act)
        ir.eval([g, gm, gf])
        res.append([read(ir, g), read(ir, gm), read(ir, gf)])
    for a, b in zip(*res):
        assert same_bits(a, b, False)
    assert np.array_equal(res[0][0], table[idx])


def test_scatter_permutation(cir, oir):
    n = 5000
    perm = np.random.default_rng(3).permutation(n).astype(np.uint32)
    vals = special_f32(np.random.default_rng(4), n)
    res = []
    for ir in (cir, oir):
        dst = ir.array_f32(np.zeros(n, np.float32))
        s = ir.scatter(ir.array_f32(vals), dst, ir.array_u32(perm))
        ir.eval([s])
        res.append(read(ir, dst))
    assert same_bits(res[0], res[1], True)


@pytest.mark.parametrize("bins", [1, 16, 65536])
def test_scatter_add_histogram_u32_bit_exact(cir, oir, bins):
    n = 200003
    rng = np.random.default_rng(bins)
    idx = rng.integers(0, bins, n).astype(np.uint32)
    table = special_u32(rng, bins)
    res = []
    for ir in (cir, oir):
        t, i = ir.array_u32(table), ir.array_u32(idx)
        b1 = ir.array_u32(np.zeros(bins, np.uint32))
        b2 = ir.array_u32(np.zeros(bins, np.uint32))
        w = ir.gather(t, i)
        s1 = ir.scatter_add(w, b1, i)                       # weighted histogram (gather + scatter-add)
        s2 = ir.scatter_add(ir.const_u32(1), b2, i, ir.lt(i, ir.const_u32(max(1, bins // 2))))  # masked count
        ir.eval([s1, s2])
        res.append([read(ir, b1), read(ir, b2)])
    for a, b in zip(*res):
        assert same_bits(a, b, False)
    assert int(res[0][1].sum()) == int((idx < max(1, bins // 2)).sum())


def test_scatter_add_f32_tolerance(cir, oir):
    n, bins = 100000, 37
    rng = np.random.default_rng(9)
    idx = rng.integers(0, bins, n).astype(np.uint32)
    vals = rng.random(n, dtype=np.float32)
    res = []
    for ir in (cir, oir):
        b = ir.array_f32(np.zeros(bins, np.float32))
        s = ir.scatter_add(ir.array_f32(vals), b, ir.array_u32(idx))
        ir.eval([s])
        res.append(read(ir, b))
    tol = 1e-6 * math.log2(n)  # order-dependent f32 accumulation, same budget as f32 sums
    assert np.all(np.abs(res[0] - res[1]) <= tol * np.abs(res[1]) + 1e-30)


# ---------------------------------------------------------------- horizontal reductions
SIZES = [1, 2, 3, 4, 5, 31, 32, 33, 511, 2048, 2049, 65537, (1 << 20) + 3]


@pytest.mark.parametrize("n", SIZES)
def test_reduce_int_bit_exact(cir, oir, n):
    data = special_u32(np.random.default_rng(n), n)
    res = []
    for ir in (cir, oir):
        u = ir.array_u32(data)
        s = ir.array_i32(data.view(np.int32))
        outs = [ir.reduce(r, v) for v in (u, s) for r in (Red.Sum, Red.Min, Red.Max)]
        res.append([read(ir, o) for o in outs])
    for a, b in zip(*res):
        assert a.shape == (1,) and same_bits(a, b, False), (n, a, b)


@pytest.mark.parametrize("n", SIZES)
def test_reduce_f32(cir, oir, n):
    rng = np.random.default_rng(n)
    for data in (rng.random(n, dtype=np.float32), (rng.random(n, dtype=np.float32) * 2 - 1).astype(np.float32)):
        res = []
        for ir in (cir, oir):
            x = ir.array_f32(data)
            res.append([read(ir, ir.reduce(r, x)) for r in (Red.Sum, Red.Min, Red.Max)])
        (cs, cmin, cmax), (os_, omin, omax) = res
        assert same_bits(cmin, omin, True) and same_bits(cmax, omax, True)
        # tolerance from BASELINE.json: 1e-6*log2(n) relative (to the sum of magnitudes for the
        # cancelling [-1,1) data set, where the result itself can be arbitrarily close to 0)
        scale = max(float(np.abs(data.astype(np.float64)).sum()), 1e-30)
        tol = 1e-6 * max(1.0, math.log2(n)) * scale
        assert abs(float(cs[0]) - float(os_[0])) <= tol, (n, cs, os_)


def test_back_to_back_reductions_keep_stream_order(cuda_backend, cir):
    """Consecutive reductions overlap through programmatic dependent launch (prims.cu: reduce_kernel).  Whatever a
    reduction reads must be complete when it reads it: results of reductions still in flight, arrays rewritten by a
    trace kernel in between, and memory someone else may write (exported pointers) all fall back to stream order."""
    rng = np.random.default_rng(7)
    n = (1 << 22) + 3
    data = [rng.integers(0, 1 << 20, n).astype(np.uint32) for _ in range(4)]
    xs = [cir.array_u32(d) for d in data]
    sums = [int(d.astype(np.uint64).sum() % (1 << 32)) for d in data]
    maxs = [int(d.max()) for d in data]
    outs = []
    for j in range(120):                          # a long chain, no host synchronisation inside
        k = j % 4
        s = cir.reduce(Red.Sum, xs[k])
        m = cir.reduce(Red.Max, xs[k])
        ss = cir.reduce(Red.Sum, s)               # input = the result of the reduction launched right before the last
        mm = cir.reduce(Red.Max, cir.reduce(Red.Min, m))   # ... and of the one launched right before
        outs.append((k, s, m, ss, mm))
    for k, s, m, ss, mm in outs:
        assert int(read(cir, s)[0]) == sums[k] and int(read(cir, m)[0]) == maxs[k]
        assert int(read(cir, ss)[0]) == sums[k] and int(read(cir, mm)[0]) == maxs[k]
    # an array rewritten by a trace kernel between two reductions of it
    y = cir.array_u32(np.zeros(n, np.uint32))
    idx = cir.arange(U32, n)
    for j in range(20):
        r0 = cir.reduce(Red.Sum, xs[0])
        sc = cir.scatter(cir.add(idx, cir.const_u32(j)), y, idx)    # y[i] = i + j
        cir.eval([sc])
        r1 = cir.reduce(Red.Sum, y)
        want = (n * (n - 1) // 2 + j * n) % (1 << 32)
        assert int(read(cir, r1)[0]) == want and int(read(cir, r0)[0]) == sums[0]
    # an exported pointer switches the overlap off for that array (others may write it); results stay right
    z = cir.array_u32(np.full(n, 3, np.uint32))
    assert cir.device_ptr(z) != 0
    for j in range(4):
        cir.reduce(Red.Sum, xs[1])
        assert int(read(cir, cir.reduce(Red.Sum, z))[0]) == 3 * n


@pytest.mark.parametrize("n", [1, 3, 4, 5, 1000, 100003, (1 << 20) + 1])
def test_fused_trace_reduce(cuda_backend, cir, oir, n):
    """An unevaluated operand is reduced by ONE generated kernel (trace + reduction epilogue); it is
    not materialised and stays unevaluated, on the device and in the oracle alike."""
    rng = np.random.default_rng(n)
    xs, ys = rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)
    res = []
    for ir in (cir, oir):
        u = ir.mul(ir.arange(U32, n), ir.const_u32(2654435761))
        s = ir.bitcast(u, I32)
        f = ir.add(ir.mul(ir.array_f32(xs), ir.array_f32(ys)), ir.const_f32(0.5))
        if ir is cir:
            cuda_backend.stats_reset()
        outs = [read(ir, ir.reduce(r, v)) for v in (u, s, f) for r in (Red.Sum, Red.Min, Red.Max)]
        if ir is cir:
            assert cuda_backend.stats()["trace_launches"] == 9      # nine fused kernels, nothing else
        assert not ir.is_buffer(u) and not ir.is_buffer(f)
        res.append(outs)
    for k, (a, b) in enumerate(zip(*res)):
        if k == 6:   # f32 sum: tolerance; everything else bit-exact
            assert abs(float(a[0]) - float(b[0])) <= 1e-6 * max(1.0, math.log2(n)) * abs(float(b[0]))
        else:
            assert same_bits(a, b, k >= 6), (n, k, a, b)


# ---------------------------------------------------------------- prefix sum / compress
SCAN_SIZES = [1, 2, 3, 4, 5, 1023, 4095, 4096, 4097, 16383, 16384, 16385, 16387, 24575, 24576, 24577, 32768, 49153,
              5 * 16384 + 3, 5 * 24576 + 3, (1 << 20) + 1, 3 * (1 << 20) + 7]


@pytest.mark.parametrize("n", SCAN_SIZES)
def test_prefix_sum_bit_exact(cir, oir, n):
    data = special_u32(np.random.default_rng(n), n)
    res = []
    for ir in (cir, oir):
        u = ir.array_u32(data)
        s = ir.array_i32(data.view(np.int32))
        outs = [ir.prefix_sum(u, True), ir.prefix_sum(u, False), ir.prefix_sum(s, True)]
        res.append([read(ir, o) for o in outs])
    for a, b in zip(*res):
        assert same_bits(a, b, False), n
    ref = np.cumsum(data.astype(np.uint64)).astype(np.uint32)
    assert np.array_equal(res[0][1], ref)


@pytest.mark.parametrize("n", SCAN_SIZES)
@pytest.mark.parametrize("density", [0.0, 0.01, 0.5, 1.0])
def test_compress_bit_exact(cir, oir, n, density):
    rng = np.random.default_rng(n + int(density * 100))
    mask = (rng.random(n) < density)
    vals = special_u32(rng, n)
    res = []
    for ir in (cir, oir):
        m = ir.array_bool(mask)
        idx, c1 = ir.compress(m)
        out, c2 = ir.compress_values(ir.array_u32(vals), m)
        res.append((c1, c2, read(ir, idx), read(ir, out)))
    assert res[0][0] == res[1][0] == res[0][1] == res[1][1] == int(mask.sum())
    assert same_bits(res[0][2], res[1][2], False) and same_bits(res[0][3], res[1][3], False)
    assert np.array_equal(res[0][2], np.nonzero(mask)[0].astype(np.uint32))
    assert np.array_equal(res[0][3], vals[mask])


def test_compress_of_traced_mask(cir, oir):
    n = 50000
    res = []
    for ir in (cir, oir):
        i = ir.arange(U32, n)
        h = ir.mul(i, ir.const_u32(747796405))
        m = ir.neq(ir.bop(Bop.And, h, ir.const_u32(4)), ir.const_u32(0))
        idx, c = ir.compress(m)
        res.append((c, read(ir, idx)))
    assert res[0][0] == res[1][0] and same_bits(res[0][1], res[1][1], False)


# ---------------------------------------------------------------- fused trace -> scan / compress
FUSED_SIZES = [1, 3, 4, 5, 4095, 4096, 4097, 8191, 12288, 12289, 24575, 24576, 24577, 49153, 5 * 24576 + 3, (1 << 20) + 1,
               3 * (1 << 20) + 7]


@pytest.mark.parametrize("n", FUSED_SIZES)
@pytest.mark.parametrize("streams", [0, 1, 2, 3, 4, 6])
def test_fused_scan_and_compress(cuda_backend, cir, oir, n, streams):
    """Unevaluated operands: ONE generated kernel evaluates the trace and scans / compacts it (scan_fused.cuh).
    `streams` arrays are staged through the TMA ring (the tile shrinks as their number grows).  Nothing is
    materialised, the operands stay unevaluated, and every word equals the oracle's."""
    rng = np.random.default_rng(1000 * streams + n)
    data = [special_u32(rng, n) for _ in range(streams)]
    res = []
    for ir in (cir, oir):
        acc = ir.mul(ir.arange(U32, n), ir.const_u32(2654435761))
        for d in data:
            acc = ir.bop(Bop.Xor, ir.add(acc, ir.array_u32(d)), ir.shr(acc, ir.const_u32(7)))
        mask = ir.neq(ir.bop(Bop.And, acc, ir.const_u32(5)), ir.const_u32(0))
        vals = ir.mul(acc, ir.const_u32(3))
        if ir is cir:
            cuda_backend.stats_reset()
        ex, inc = ir.prefix_sum(acc, True), ir.prefix_sum(acc, False)
        idx, c1 = ir.compress(mask)
        out, c2 = ir.compress_values(vals, mask)
        if ir is cir:
            st = cuda_backend.stats()
            assert st["trace_launches"] == 4, st          # four fused kernels and nothing else
        assert not ir.is_buffer(acc) and not ir.is_buffer(mask) and not ir.is_buffer(vals)
        res.append((c1, c2, read(ir, ex), read(ir, inc), read(ir, idx), read(ir, out), ir.as_slice_eval(acc, U32)))
    assert res[0][0] == res[1][0] and res[0][1] == res[1][1] == res[0][0]
    for k in range(2, 7):
        assert same_bits(res[0][k], res[1][k], False), (n, streams, k)
    acc_np = res[1][6]
    assert np.array_equal(res[0][3], np.cumsum(acc_np.astype(np.uint64)).astype(np.uint32))
    assert np.array_equal(res[0][4], np.nonzero(acc_np & 5)[0].astype(np.uint32))


@pytest.mark.parametrize("n", FUSED_SIZES)
def test_lagged_fused_kernels_small_traces(cuda_backend, cir, oir, n):
    """Traces of a few nodes over at most one streamed array take the LAGGED fused kernels (look-back one tile behind,
    16384- / 12288-lane tiles): all four modes, with and without a streamed array, every tile-boundary size."""
    x = special_u32(np.random.default_rng(n), n)
    res = []
    for ir in (cir, oir):
        xv = ir.array_u32(x)
        lanes = ir.arange(U32, n)
        m_x, m_l = ir.gt(xv, ir.const_u32(1 << 31)), ir.eq(ir.bop(Bop.And, lanes, ir.const_u32(3)), ir.const_u32(1))
        sh = ir.shr(xv, ir.const_u32(3))
        if ir is cir:
            cuda_backend.stats_reset()
        outs = [ir.prefix_sum(sh, True), ir.prefix_sum(sh, False), ir.prefix_sum(ir.mul(lanes, ir.const_u32(5)), True)]
        (i1, c1), (v1, c2) = ir.compress(m_x), ir.compress_values(xv, m_x)
        (i2, c3), (v2, c4) = ir.compress(m_l), ir.compress_values(ir.mul(lanes, ir.const_u32(7)), m_l)
        if ir is cir:
            assert cuda_backend.stats()["trace_launches"] == 7
        res.append(([read(ir, o) for o in outs + [i1, v1, i2, v2]], (c1, c2, c3, c4)))
    assert res[0][1] == res[1][1]
    for a, b in zip(res[0][0], res[1][0]):
        assert same_bits(a, b, False), n
    assert np.array_equal(res[0][0][4], x[x > (1 << 31)])


def test_fused_compress_values_of_a_bound_array(cuda_backend, cir, oir):
    """compress_values(x, x > t): the C28 "fused-mask" shape — x is streamed once (4 B/lane), the mask never exists."""
    n = 5 * 24576 + 1234
    x = np.random.default_rng(5).random(n, dtype=np.float32)
    res = []
    for ir in (cir, oir):
        xv = ir.array_f32(x)
        m = ir.gt(xv, ir.const_f32(0.75))
        if ir is cir:
            cuda_backend.stats_reset()
        out, c = ir.compress_values(xv, m)
        if ir is cir:
            assert cuda_backend.stats()["trace_launches"] == 1
        res.append((c, read(ir, out)))
    assert res[0][0] == res[1][0] == int((x > 0.75).sum())
    assert same_bits(res[0][1], res[1][1], True) and np.array_equal(res[0][1], x[x > 0.75])


def test_fused_scan_with_gather_and_fallbacks(cuda_backend, cir, oir, monkeypatch):
    """A gather inside the scanned trace (pointer parameter next to the streamed arrays); traces the fused kernel
    does not take (side effects, > 6 streamed arrays) go through temporaries and give the same words."""
    n, m = 70001, 997
    rng = np.random.default_rng(11)
    table, idx = special_u32(rng, m), rng.integers(0, m, n).astype(np.uint32)
    many = [rng.integers(0, 1 << 20, n).astype(np.uint32) for _ in range(8)]
    res = []
    for ir in (cir, oir):
        g = ir.gather(ir.array_u32(table), ir.array_u32(idx))
        a = ir.prefix_sum(g, True)
        wide = ir.array_u32(many[0])
        for d in many[1:]:
            wide = ir.add(wide, ir.array_u32(d))
        b = ir.prefix_sum(wide, False)                         # 8 streamed arrays: not fused
        tgt = ir.array_u32(np.zeros(n, np.uint32))
        sc = ir.scatter(ir.arange(U32, n), tgt, ir.arange(U32, n))   # value of a scatter var = its source
        c, cnt = ir.compress(ir.neq(ir.bop(Bop.And, sc, ir.const_u32(1)), ir.const_u32(0)))
        assert not ir.is_buffer(g) and not ir.is_buffer(wide) and not ir.is_buffer(sc)
        res.append((read(ir, a), read(ir, b), read(ir, c), cnt, read(ir, tgt)))
    for k in (0, 1, 2, 4):
        assert same_bits(res[0][k], res[1][k], False), k
    assert res[0][3] == res[1][3] == n // 2


# ---------------------------------------------------------------- errors (reference panics -> status)
def test_errors_match_oracle(cir, oir):
    from vkjit_b200 import VkjitError, VkjitSizeError, VkjitTypeError
    for ir in (cir, oir):
        a, b = ir.array_f32([1, 2, 3]), ir.array_f32([1, 2])
        with pytest.raises(VkjitSizeError):
            ir.eval([ir.add(a, b)])                      # internal.rs:699-702
        with pytest.raises(VkjitSizeError):
            ir.eval([ir.add(ir.const_f32(1), ir.const_f32(2))])  # internal.rs:1202 num.unwrap()
        with pytest.raises(VkjitTypeError):
            ir.select(ir.lt(a, a), a, ir.array_u32([1, 2, 3]))   # internal.rs:232
        z = ir.add(a, a)
        ir.eval([z])
        with pytest.raises(VkjitTypeError):
            ir.as_slice(z, U32)                                  # internal.rs:447
        with pytest.raises(VkjitError):
            ir.eval([ir.gather(ir.add(a, a), ir.arange(U32, 3))])  # internal.rs:1054
        with pytest.raises(VkjitError):
            ir.eval([ir.scatter(a, ir.const_f32(0), ir.arange(U32, 3))])  # internal.rs:1059-1062
        # the Ir stays usable after an error
        ir.eval([ir.add(a, a)])


def test_disk_cubin_cache(cuda_backend, cir, tmp_path, monkeypatch):
    """$VKJIT_CACHE_DIR: a kernel-cache miss whose cubin is on disk skips NVRTC."""
    monkeypatch.setenv("VKJIT_CACHE_DIR", str(tmp_path))
    def run():
        x = cir.add(cir.mul(cir.arange(F32, 777), cir.const_f32(1.25)), cir.const_f32(-3.0))
        cir.eval([x])
        return read(cir, x)
    cuda_backend.cache_clear(); cuda_backend.stats_reset()
    a = run()
    assert cuda_backend.stats()["cache_misses"] == 1 and cuda_backend.stats()["disk_hits"] == 0
    assert len(list(tmp_path.glob("*.cubin"))) == 1
    cuda_backend.cache_clear(); cuda_backend.stats_reset()
    b = run()
    st = cuda_backend.stats()
    assert st["cache_misses"] == 1 and st["disk_hits"] == 1
    assert same_bits(a, b, True)


def test_mixed_size_schedule_splits_into_kernels(cuda_backend, cir, oir):
    """one eval with roots of different sizes runs one kernel per size group (the reference
    asserts, internal.rs:697-706); a conflict inside a single root's expression is still an error."""
    from vkjit_b200 import VkjitSizeError
    res = []
    for ir in (cir, oir):
        a = ir.add(ir.arange(U32, 1000), ir.const_u32(1))
        b = ir.mul(ir.arange(F32, 77), ir.const_f32(0.5))
        c = ir.bop(Bop.Xor, ir.arange(U32, 1000), ir.const_u32(0xFFFF))
        d = ir.add(ir.arange(I32, 5), ir.const_i32(-2))
        if ir is cir:
            cuda_backend.stats_reset()
        ir.eval([a, b, c, d])
        if ir is cir:
            assert cuda_backend.stats()["trace_launches"] == 3       # sizes 1000 (a, c), 77, 5
        res.append([read(ir, v) for v in (a, b, c, d)])
        with pytest.raises(VkjitSizeError):
            ir.eval([ir.add(ir.arange(U32, 3), ir.arange(U32, 4)), ir.arange(U32, 9)])
        ir.eval([ir.arange(U32, 9)])                                 # still usable
    for x, y in zip(*res):
        assert x.tobytes() == y.tobytes()


@pytest.mark.parametrize("bins", [16, 65536, 100000])
def test_scatter_add_privatised_in_shared_memory(cuda_backend, cir, oir, bins):
    """Launches of >= 2^22 lanes keep up to 49152 bins of the scatter_add target in shared memory (the rest
    go to L2 as before) and flush them once per CTA: u32 bit-exact, f32 within the sum tolerance."""
    n = (1 << 22) + 5
    rng = np.random.default_rng(bins)
    idx = rng.integers(0, bins, n).astype(np.uint32)
    vals = rng.random(n, dtype=np.float32)
    res = []
    for ir in (cir, oir):
        i = ir.array_u32(idx)
        b1 = ir.array_u32(np.full(bins, 7, np.uint32))           # non-zero start: the flush must ADD
        b2 = ir.array_f32(np.zeros(bins, np.float32))
        s1 = ir.scatter_add(ir.mul(i, ir.const_u32(2654435761)), b1, i, ir.neq(ir.bop(Bop.And, i, ir.const_u32(3)), ir.const_u32(0)))
        s2 = ir.scatter_add(ir.array_f32(vals), b2, i)
        ir.eval([s1])
        ir.eval([s2])
        res.append((read(ir, b1), read(ir, b2)))
    assert same_bits(res[0][0], res[1][0], False)
    tol = 1e-6 * 22
    assert np.all(np.abs(res[0][1] - res[1][1]) <= tol * np.abs(res[1][1]) + 1e-30)


def test_classic_scan_kernels_still_agree(tmp_path):
    """`VKJIT_SCAN_IMPL=classic` selects the immediate-look-back kernels that the lagged ones replaced (kept for A/B
    measurements, profiles/r01_scan_history.md).  The switch is read once per process, hence the subprocess."""
    import subprocess
    import sys
    script = tmp_path / "classic.py"
    script.write_text('''
import sys
sys.path.insert(0, %r)
import numpy as np
import vkjit_b200 as vk
from vkjit_b200.ir import Ir, VarType as T
vk.init(0)
ir = Ir()
for n in (5, 24577, 3 * 24576 + 11, (1 << 21) + 3):
    rng = np.random.default_rng(n)
    x = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    m = rng.random(n) < 0.4
    xv, mv = ir.array_u32(x), ir.array_bool(m)
    ex = ir.as_slice(ir.prefix_sum(xv, True), T.U32)
    ref = np.cumsum(x.astype(np.uint64)).astype(np.uint32)
    assert ex[0] == 0 and np.array_equal(ex[1:], ref[:-1]), n
    idx, c1 = ir.compress(mv)
    val, c2 = ir.compress_values(xv, mv)
    assert c1 == c2 == int(m.sum())
    assert np.array_equal(ir.as_slice(idx, T.U32), np.nonzero(m)[0].astype(np.uint32))
    assert np.array_equal(ir.as_slice(val, T.U32), x[m])
    fm = ir.gt(xv, ir.const_u32(1 << 31))                      # fused, immediate look-back
    fv, c3 = ir.compress_values(xv, fm)
    assert c3 == int((x > (1 << 31)).sum()) and np.array_equal(ir.as_slice(fv, T.U32), x[x > (1 << 31)])
print("classic ok")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    env = dict(os.environ, VKJIT_SCAN_IMPL="classic")
    r = subprocess.run([sys.executable, str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "classic ok" in r.stdout, r.stdout + r.stderr


def test_side_effect_of_a_primitives_operand_runs_once(cir, oir):
    """An eager primitive evaluates an unevaluated operand inside its own kernel WITHOUT committing it — unless the
    operand's trace scatters: then it is committed first, so the scatter runs once, not again at the var's own eval
    (round-1 defect: dst ended up [2 2 2 2])."""
    n = 4096
    for ir in (cir, oir):
        for prim in ("reduce", "prefix_sum", "compress"):
            dst = ir.array_u32(np.zeros(n, np.uint32))
            ones = ir.add(ir.arange(U32, n), ir.const_u32(1))
            s = ir.scatter_add(ones, dst, ir.arange(U32, n))
            if prim == "reduce":
                r = ir.reduce(Red.Sum, s)
                assert int(ir.as_slice(r, U32)[0]) == n * (n + 1) // 2
            elif prim == "prefix_sum":
                r = ir.prefix_sum(s, False)
                assert int(ir.as_slice(r, U32)[-1]) == n * (n + 1) // 2
            else:
                r, cnt = ir.compress_values(s, ir.lt(s, ir.const_u32(10)))
                assert cnt == 9
            assert ir.is_buffer(s)                       # committed by the primitive
            ir.eval([s])                                 # a Binding root: copied, nothing re-executed
            assert np.array_equal(ir.as_slice(dst, U32), np.arange(1, n + 1, dtype=np.uint32)), prim


@pytest.mark.parametrize("n", [1, 16383, 16384, (2 << 20) // 4 + 1, (16 << 20) // 4 - 1, (16 << 20) // 4, (18 << 20) // 4 + 5, 9 * (1 << 20) + 3])
def test_upload_and_readback_through_the_staging_ring(cir, n):
    """Pageable host memory crosses PCIe through the pinned staging ring (csrc/staging.cpp: 4 x 8 MiB chunks, threaded
    memcpy) from 16 MiB on, the driver's pageable path below; sizes around the pinned-pool threshold of the front-end
    (64 KiB), around the ring threshold, and several chunks with a ragged tail.  The
    upload returns once the caller's buffer is copied out, so overwriting it right away must not change the array."""
    rng = np.random.default_rng(n)
    src = rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32)
    keep = src.copy()
    v = cir.array_u32(src)
    src[:] = 0xDEADBEEF                                   # the reference's contract: the slice may be reused at once
    y = cir.add(v, cir.const_u32(1))
    got = cir.as_slice_eval(y, U32)                       # result array from the pinned pool (>= 64 KiB) or numpy
    assert np.array_equal(got, keep + np.uint32(1))
    host = np.empty(n, np.uint32)                         # the C ABI into the caller's pageable buffer
    cir.read_into(v, U32, host.ctypes.data, host.nbytes)
    assert np.array_equal(host, keep)
    again = cir.as_slice(v, U32)                          # a second read: another block of the pool, same contents
    assert np.array_equal(again, keep) and (n * 4 < (64 << 10) or again.ctypes.data != got.ctypes.data)


@pytest.mark.parametrize("variant", ["default", "ctrl", "ctrl_deep", "ctrl_small_tiles", "wreg", "early", "lag_packed"])
def test_fused_compress_kernel_variants(tmp_path, variant):
    """The fused trace -> compress kernels in every schedule the library can generate (scan_fused.cuh): the shipped lagged
    kernel, the opt-in control-warp pipeline (VKJIT_SCAN_CTRL=1: a dedicated warp owns the totals scan / publish / anchored
    look-back, the workers write tile k - D; lane-by-lane and coalesced warp-row output) at several lags and tile sizes, the status window
    through strong loads (VKJIT_SCAN_WREG=1), requested early (VKJIT_SCAN_EARLY=1), and the packed anchored look-back in the
    lagged kernel (VKJIT_LAG_PACKED=1) — all bit-exact against the oracle:
    indices and values, a mask computed from the streamed array and one computed from the lane index, sizes from one
    lane over ragged tiles to several generations of the persistent grid.  The switches are read once per process."""
    import subprocess
    import sys
    script = tmp_path / "fc.py"
    script.write_text('''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import vkjit_b200 as vk
from oracle_lib import OracleIr
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
def h(ir, x, seed):
    s = ir.add(ir.mul(ir.bop(Bop.Xor, x, ir.const_u32(seed)), ir.const_u32(747796405)), ir.const_u32(2891336453))
    return ir.bop(Bop.Xor, ir.bop(Bop.Shr, s, ir.const_u32(15)), s)
for n in (1, 5, 4095, 8191, 8192, 8193, 100003, 296 * 8192 + 17, 3 * 296 * 8192 + 4099, 9 * 296 * 4096 + 1):
    res = []
    for ir in (Ir(), OracleIr()):
        lanes = ir.arange(T.U32, n)
        vals = h(ir, lanes, 3)
        ir.eval([vals])
        got = []
        for thr in (0x80000000, 0xF0000000, 0x00000010):            # p = 1/2, 1/16, ~1
            m = ir.gt(vals, ir.const_u32(thr))
            r, k = ir.compress_values(vals, m); got.append((k, ir.as_slice(r, T.U32).copy()))
            m = ir.gt(vals, ir.const_u32(thr))
            r, k = ir.compress(m); got.append((k, ir.as_slice(r, T.U32).copy()))
        m = ir.neq(ir.bop(Bop.And, h(ir, lanes, 4), ir.const_u32(1)), ir.const_u32(0))   # mask from the lane index
        r, k = ir.compress_values(vals, m); got.append((k, ir.as_slice(r, T.U32).copy()))
        m = ir.lt(lanes, ir.const_u32(0))                             # nothing selected
        r, k = ir.compress(m); got.append((k, ir.as_slice(r, T.U32).copy()))
        res.append(got)
        ir.close()
    for (ka, a), (kb, b) in zip(*res):
        assert ka == kb and np.array_equal(a, b), (n, ka, kb)
print("fused compress ok")
''' % (ROOT, os.path.join(ROOT, "tests")))
    env = {"default": {}, "ctrl": {"VKJIT_SCAN_CTRL": "1"},
           "ctrl_deep": {"VKJIT_SCAN_CTRL": "1", "VKJIT_CTRL_LAG": "4", "VKJIT_CTRL_DEPTH": "6", "VKJIT_FSCAN_DIAG": "2"},
           "ctrl_small_tiles": {"VKJIT_SCAN_CTRL": "1", "VKJIT_CTRL_VPT": "2", "VKJIT_CTRL_LAG": "1", "VKJIT_CTRL_DEPTH": "1"},
           "wreg": {"VKJIT_SCAN_WREG": "1"}, "early": {"VKJIT_SCAN_EARLY": "1"}, "lag_packed": {"VKJIT_LAG_PACKED": "1"}}[variant]
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "fused compress ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


@pytest.mark.parametrize("pass_kb", ["128", "96", "40"])
def test_scatter_add_bin_range_passes(tmp_path, pass_kb):
    """Opt-in bin-range passes (VKJIT_SADD_PASSES=1, program.cpp variant 3): a trace whose ONE side effect is a scatter_add
    into a target too large for one CTA's shared memory runs in P launches, pass p executing only the lanes whose bin lies
    in its range — every lane's roots are written exactly once, every atomic lands in shared memory.  Bit-exact against
    the oracle: gather + scatter_add with the same index (H26's shape), a masked I32 scatter_add with a root that is NOT
    the scatter var, an f32 target (exactly representable sums), bins that do not divide into the passes, and a target so
    large that the pass variant is refused (> 4 passes: the per-CTA variant runs instead)."""
    import subprocess
    import sys
    script = tmp_path / "passes.py"
    script.write_text('''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import vkjit_b200 as vk
from oracle_lib import OracleIr
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
for bins, n in [(65536, (1 << 22) + 5), (70001, (1 << 22) + 64), (40000, (1 << 22) + 1), (300007, (1 << 22) + 3)]:
    out = []
    for ir in (Ir(), OracleIr()):
        lanes = ir.arange(T.U32, n)
        h = ir.mul(ir.bop(Bop.Xor, lanes, ir.const_u32(0x9E3779B9)), ir.const_u32(747796405))
        idx = ir.bop(Bop.Shr, h, ir.const_u32(9))
        idx = ir.sub(idx, ir.mul(ir.div(idx, ir.const_u32(bins)), ir.const_u32(bins)))        # idx mod bins
        ir.eval([idx])
        table = ir.array_u32((np.arange(bins, dtype=np.uint64) * 2654435761 %% (1 << 32)).astype(np.uint32))
        du = ir.array_u32(np.zeros(bins, np.uint32))
        s1 = ir.scatter_add(ir.gather(table, idx), du, idx)                  # H26: gather and scatter with the same index
        ir.eval([s1])
        di = ir.array_i32(np.zeros(bins, np.int32))
        wi = ir.sub(ir.cast(ir.bop(Bop.And, h, ir.const_u32(1023)), T.I32), ir.const_i32(700))
        mask = ir.neq(ir.bop(Bop.And, h, ir.const_u32(4)), ir.const_u32(0))
        s2 = ir.scatter_add(wi, di, idx, mask)
        other = ir.add(ir.bop(Bop.Shr, h, ir.const_u32(7)), idx)            # a second root next to the scatter var
        ir.eval([s2, other])
        df = ir.array_f32(np.zeros(bins, np.float32))
        s3 = ir.scatter_add(ir.cast(ir.bop(Bop.And, h, ir.const_u32(3)), T.F32), df, idx)
        ir.eval([s3])
        out.append([ir.as_slice(du, T.U32).copy(), ir.as_slice(s1, T.U32).copy(), ir.as_slice(di, T.I32).copy(), ir.as_slice(s2, T.I32).copy(),
                    ir.as_slice(other, T.U32).copy(), ir.as_slice(df, T.F32).copy(), ir.as_slice(s3, T.F32).copy()])
        ir.close()
    for k, (a, b) in enumerate(zip(*out)):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), (bins, n, k)
print("passes ok")
''' % (ROOT, os.path.join(ROOT, "tests")))
    env = {"VKJIT_SADD_PASSES": "1", "VKJIT_PASS_KB": pass_kb}
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "passes ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]


@pytest.mark.parametrize("variant", ["plain", "agg", "cluster"])
def test_scatter_add_hot_bins(tmp_path, variant):
    """Integer scatter_add when the lanes of a warp collide (few distinct bins), plain and with the opt-in warp-aggregated
    path (VKJIT_AGG=1: every warp probes once, then one atomic per distinct bin through match.any + redux.sync,
    program.cpp: kSaddHelper) — bit-exact against the oracle, with and without a mask (lanes that join a warp's
    scatter_add later), U32 and I32 (negative values), small launches and the shared-memory-privatised variant of launches
    >= 2^22 lanes — per CTA, and (VKJIT_SADD_CLUSTER=1) split over the two CTAs of a thread-block cluster through
    distributed shared memory for targets that do not fit one CTA.  The switches are read once per process, hence the
    subprocess."""
    import subprocess
    import sys
    script = tmp_path / "hot.py"
    script.write_text('''
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
import vkjit_b200 as vk
from oracle_lib import OracleIr
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
for bins, n in [(1, 4099), (4, 100003), (16, (1 << 22) + 5), (1000, 50001), (65536, (1 << 22) + 1), (70000, (1 << 22) + 7), (120001, (1 << 22) + 64)]:
    out = []
    for ir in (Ir(), OracleIr()):
        lanes = ir.arange(T.U32, n)
        h = ir.mul(ir.bop(Bop.Xor, lanes, ir.const_u32(0x9E3779B9)), ir.const_u32(747796405))
        idx = ir.bop(Bop.Shr, h, ir.const_u32(9))
        idx = ir.sub(idx, ir.mul(ir.div(idx, ir.const_u32(bins)), ir.const_u32(bins)))        # idx mod bins
        du = ir.array_u32(np.zeros(bins, np.uint32))
        di = ir.array_i32(np.zeros(bins, np.int32))
        wi = ir.sub(ir.cast(ir.bop(Bop.And, h, ir.const_u32(1023)), T.I32), ir.const_i32(700))  # in [-700, 323]
        mask = ir.neq(ir.bop(Bop.And, h, ir.const_u32(4)), ir.const_u32(0))
        s1 = ir.scatter_add(ir.bop(Bop.Shr, h, ir.const_u32(3)), du, idx)
        s2 = ir.scatter_add(wi, di, idx, mask)
        ir.eval([s1, s2])
        out.append((ir.as_slice(du, T.U32).copy(), ir.as_slice(di, T.I32).copy()))
        ir.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]), (bins, n)
print("hot bins ok")
''' % (ROOT, os.path.join(ROOT, "tests")))
    env = {"plain": {}, "agg": {"VKJIT_AGG": "1"}, "cluster": {"VKJIT_SADD_CLUSTER": "1"}}[variant]
    r = subprocess.run([sys.executable, str(script)], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "hot bins ok" in r.stdout, r.stdout[-1000:] + r.stderr[-3000:]
