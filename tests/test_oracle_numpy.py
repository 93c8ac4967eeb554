"""Independent cross-check of the CPU oracle against NumPy for the operations the reference does not
pin ("parity unpinned"): IEEE f32 arithmetic, modular integer arithmetic, comparisons,
casts, select, gather/scatter, bit ops, reductions, prefix sum, compress.  NumPy is a second,
unrelated implementation of the same published semantics (IEEE-754 binary32, two's complement)."""
import numpy as np
import pytest

from trace_gen import special_f32, special_u32
from vkjit_b200.ir import Bop, Red, Uop, VarType as T

F32, U32, I32, BOOL = T.F32, T.U32, T.I32, T.Bool
N = 4099


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def eq_f32(a, b):
    return bool(np.all((bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture()
def data():
    rng = np.random.default_rng(42)
    return dict(fa=special_f32(rng, N), fb=special_f32(rng, N), ua=special_u32(rng, N), ub=special_u32(rng, N))


def test_f32_arithmetic_is_ieee(oir, data):
    a, b = oir.array_f32(data["fa"]), oir.array_f32(data["fb"])
    with np.errstate(all="ignore"):
        exp = {Bop.Add: data["fa"] + data["fb"], Bop.Sub: data["fa"] - data["fb"], Bop.Mul: data["fa"] * data["fb"],
               Bop.Div: data["fa"] / data["fb"]}
    for k, e in exp.items():
        assert eq_f32(oir.as_slice_eval(oir.bop(k, a, b), F32), e.astype(np.float32)), k
    with np.errstate(all="ignore"):
        assert eq_f32(oir.as_slice_eval(oir.sqrt(a), F32), np.sqrt(data["fa"]))
    # x*y+c is two roundings (OpFMul, OpFAdd), never an FMA
    z = oir.add(oir.mul(a, b), oir.const_f32(0.5))
    with np.errstate(all="ignore"):
        assert eq_f32(oir.as_slice_eval(z, F32), (data["fa"] * data["fb"]).astype(np.float32) + np.float32(0.5))


def test_integer_arithmetic_is_modular(oir, data):
    ua, ub = data["ua"], data["ub"]
    a, b = oir.array_u32(ua), oir.array_u32(ub)
    ia, ib = oir.array_i32(ua.view(np.int32)), oir.array_i32(ub.view(np.int32))
    with np.errstate(over="ignore"):
        for k, e in ((Bop.Add, ua + ub), (Bop.Sub, ua - ub), (Bop.Mul, ua * ub), (Bop.And, ua & ub), (Bop.Or, ua | ub), (Bop.Xor, ua ^ ub),
                     (Bop.Shl, ua << (ub & 31)), (Bop.Shr, ua >> (ub & 31)), (Bop.Min, np.minimum(ua, ub)), (Bop.Max, np.maximum(ua, ub))):
            assert np.array_equal(oir.as_slice_eval(oir.bop(k, a, b), U32), e.astype(np.uint32)), k
        sa, sb = ua.view(np.int32), ub.view(np.int32)
        for k, e in ((Bop.Add, sa + sb), (Bop.Mul, sa * sb), (Bop.Shr, sa >> (ub & 31).astype(np.int32)), (Bop.Min, np.minimum(sa, sb))):
            assert np.array_equal(oir.as_slice_eval(oir.bop(k, ia, ib), I32), e.astype(np.int32)), k
    d = oir.const_u32(7)
    assert np.array_equal(oir.as_slice_eval(oir.div(a, d), U32), ua // 7)
    di = oir.const_i32(7)
    assert np.array_equal(oir.as_slice_eval(oir.div(ia, di), I32), np.trunc(sa / 7.0).astype(np.int32))   # OpSDiv truncates toward 0
    assert np.array_equal(oir.as_slice_eval(oir.uop(Uop.Not, a), U32), ~ua)
    assert np.array_equal(oir.as_slice_eval(oir.uop(Uop.Neg, ia), I32), (-sa.astype(np.int64)).astype(np.int32))


def test_comparisons_casts_select(oir, data):
    fa, fb, ua = data["fa"], data["fb"], data["ua"]
    a, b, u = oir.array_f32(fa), oir.array_f32(fb), oir.array_u32(ua)
    with np.errstate(all="ignore"):
        for k, e in ((Bop.Lt, fa < fb), (Bop.Gt, fa > fb), (Bop.Leq, fa <= fb), (Bop.Geq, fa >= fb), (Bop.Eq, fa == fb),
                     (Bop.Neq, (fa < fb) | (fa > fb))):            # FOrdNotEqual: false on NaN
            assert np.array_equal(oir.as_slice_eval(oir.bop(k, a, b), BOOL).astype(bool), e), k
    assert eq_f32(oir.as_slice_eval(oir.cast(u, F32), F32), ua.astype(np.float32))                       # ConvertUToF, RNE
    assert eq_f32(oir.as_slice_eval(oir.cast(oir.bitcast(u, I32), F32), F32), ua.view(np.int32).astype(np.float32))
    with np.errstate(all="ignore"):
        d = fa.astype(np.float64)
        f2u = np.where(np.isnan(d) | (d <= 0), 0.0, np.where(d >= 4294967296.0, 4294967295.0, np.trunc(d))).astype(np.uint32)
    assert np.array_equal(oir.as_slice_eval(oir.cast(a, U32), U32), f2u)                                 # round toward zero, saturating
    sel = oir.select(oir.lt(a, b), a, b)
    with np.errstate(all="ignore"):
        assert eq_f32(oir.as_slice_eval(sel, F32), np.where(fa < fb, fa, fb))


def test_gather_scatter_scatter_add(oir, data):
    rng = np.random.default_rng(1)
    ua = data["ua"]
    idx = rng.integers(0, N, N).astype(np.uint32)
    t, i = oir.array_u32(ua), oir.array_u32(idx)
    assert np.array_equal(oir.as_slice_eval(oir.gather(t, i), U32), ua[idx])
    act = (idx % 3 == 0)
    g = oir.gather(t, i, oir.array_bool(act))
    assert np.array_equal(oir.as_slice_eval(g, U32), np.where(act, ua[idx], 0))
    perm = rng.permutation(N).astype(np.uint32)
    dst = oir.array_u32(np.zeros(N, np.uint32))
    oir.eval([oir.scatter(t, dst, oir.array_u32(perm))])
    exp = np.zeros(N, np.uint32); exp[perm] = ua
    assert np.array_equal(oir.as_slice(dst, U32), exp)
    bins = oir.array_u32(np.zeros(64, np.uint32))
    oir.eval([oir.scatter_add(t, bins, oir.array_u32(idx % 64))])
    e = np.zeros(64, np.uint64); np.add.at(e, idx % 64, ua.astype(np.uint64))
    assert np.array_equal(oir.as_slice(bins, U32), e.astype(np.uint32))


@pytest.mark.parametrize("threads", [1, 4])
def test_reduce_scan_compress(oir, oracle_api, data, threads):
    oracle_api.call("set_threads", threads)
    try:
        n = 300001
        rng = np.random.default_rng(2)
        ua = special_u32(rng, n)
        fa = rng.random(n, dtype=np.float32)
        u, f = oir.array_u32(ua), oir.array_f32(fa)
        assert oir.as_slice(oir.reduce(Red.Sum, u), U32)[0] == np.uint32(ua.astype(np.uint64).sum() & 0xFFFFFFFF)
        assert oir.as_slice(oir.reduce(Red.Min, u), U32)[0] == ua.min() and oir.as_slice(oir.reduce(Red.Max, u), U32)[0] == ua.max()
        s = ua.view(np.int32)
        si = oir.array_i32(s)
        assert oir.as_slice(oir.reduce(Red.Min, si), I32)[0] == s.min() and oir.as_slice(oir.reduce(Red.Max, si), I32)[0] == s.max()
        assert oir.as_slice(oir.reduce(Red.Sum, f), F32)[0] == np.float32(fa.astype(np.float64).sum())   # f64 accumulation
        assert oir.as_slice(oir.reduce(Red.Max, f), F32)[0] == fa.max()
        c = np.cumsum(ua.astype(np.uint64)).astype(np.uint32)
        assert np.array_equal(oir.as_slice(oir.prefix_sum(u, False), U32), c)
        assert np.array_equal(oir.as_slice(oir.prefix_sum(u, True), U32), np.concatenate([[0], c[:-1]]).astype(np.uint32))
        mask = (ua & 4) != 0
        m = oir.array_bool(mask)
        idx, k = oir.compress(m)
        vals, k2 = oir.compress_values(u, m)
        assert k == k2 == int(mask.sum())
        assert np.array_equal(oir.as_slice(idx, U32), np.nonzero(mask)[0].astype(np.uint32))
        assert np.array_equal(oir.as_slice(vals, U32), ua[mask])
    finally:
        oracle_api.call("set_threads", 1)


def test_transcendentals_within_one_ulp_of_numpy_f64(oir):
    from trace_gen import ulp_diff
    x = np.random.default_rng(3).uniform(-30, 30, 5000).astype(np.float32)
    p = np.random.default_rng(4).uniform(1e-6, 1e6, 5000).astype(np.float32)
    a, b = oir.array_f32(x), oir.array_f32(p)
    for got, exp in ((oir.exp(a), np.exp(x.astype(np.float64))), (oir.log(b), np.log(p.astype(np.float64))),
                     (oir.sin(a), np.sin(x.astype(np.float64))), (oir.cos(a), np.cos(x.astype(np.float64)))):
        assert int(ulp_diff(oir.as_slice_eval(got, F32), exp.astype(np.float32)).max()) <= 1


def test_hash_generator_matches_numpy(oracle_api):
    import ctypes as C
    n, seed = 10007, 0xB2000021
    out = np.empty(n, np.uint32)
    oracle_api.call("fill_hash", out.ctypes.data_as(C.c_void_p), n, 5, seed, 0)
    with np.errstate(over="ignore"):
        x = (np.arange(5, n + 5, dtype=np.uint32) ^ np.uint32(seed))
        s = x * np.uint32(747796405) + np.uint32(2891336453)
        w = ((s >> ((s >> np.uint32(28)) + np.uint32(4))) ^ s) * np.uint32(277803737)
        h = (w >> np.uint32(22)) ^ w
    assert np.array_equal(out, h)
    f = np.empty(n, np.float32)
    oracle_api.call("fill_hash", f.ctypes.data_as(C.c_void_p), n, 5, seed, 1)
    assert np.array_equal(f, (h >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24))
