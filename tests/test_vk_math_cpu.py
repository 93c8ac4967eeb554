"""vk_math.h on the CPU (through the oracle, which compiles it): the committed known-answer vectors, the 1-ulp bound
against independently computed correctly rounded values, and a 2^20-input sweep against NumPy's f64 libm.
The exhaustive 2^32-input record is profiles/r02_vk_math_ulp.md (tools/vk_math_ulp.cpp)."""
import json
import os

import numpy as np

from oracle_lib import OracleIr
from trace_gen import ulp_diff
from vkjit_b200.ir import VarType as T

HERE = os.path.dirname(os.path.abspath(__file__))


def run(xs):
    o = OracleIr()
    x = o.array_f32(xs)
    r = [o.exp(x), o.log(x), o.sin(x), o.cos(x)]
    o.eval(r)
    out = [o.as_slice(v, T.F32).copy() for v in r]
    o.close()
    return dict(zip(("exp", "log", "sin", "cos"), out))


def test_known_answers_and_one_ulp_of_correctly_rounded():
    g = json.load(open(os.path.join(HERE, "golden", "vk_math_golden.json")))
    xs = np.array(g["inputs_bits"], dtype=np.uint32).view(np.float32)
    got = run(xs)
    for name, cols in g["functions"].items():
        mine = got[name].view(np.uint32)
        pinned = np.array(cols["vk_math_bits"], dtype=np.uint32)
        assert np.array_equal(mine, pinned), (name, xs[mine != pinned][:4])          # same bits on every box / compiler
        exact = np.array(cols["correctly_rounded_bits"], dtype=np.uint32).view(np.float32)
        d = ulp_diff(got[name], exact)
        inf_mismatch = np.isinf(exact) != np.isinf(got[name])
        assert int(d[~inf_mismatch].max()) <= 1 and not inf_mismatch.any(), (name, int(d.max()), xs[np.argmax(d)])
        assert np.array_equal(np.isnan(got[name]), np.isnan(exact)), name
        assert (got[name].view(np.uint32)[np.isnan(exact)] == 0x7FC00000).all()      # the one NaN pattern


def test_sweep_against_f64_libm():
    rng = np.random.default_rng(7)
    n = 1 << 18
    xs = np.concatenate([rng.uniform(-100, 100, n), rng.uniform(-2.0 ** 20, 2.0 ** 20, n),
                         rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-44, 38, n),
                         rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32).view(np.float32).astype(np.float64)]).astype(np.float32)
    xs = xs[np.isfinite(xs)]
    got = run(xs)
    x64 = xs.astype(np.float64)
    with np.errstate(all="ignore"):
        want = {"exp": np.exp(x64), "log": np.log(x64), "sin": np.sin(x64), "cos": np.cos(x64)}
    for name, w in want.items():
        w32 = w.astype(np.float32)
        ok = np.isfinite(w32) & (w32 != 0)
        d = ulp_diff(got[name][ok], w32[ok])
        assert int(d.max()) <= 1, (name, int(d.max()), xs[ok][np.argmax(d)])
        assert np.array_equal(np.isnan(got[name]), np.isnan(w32)), name
