"""Random build / eval / drop sequences on the CUDA path: results equal the oracle's, dropped handles never
break later evaluations, and once every handle is released the device pool is back at its baseline
(no leaked arrays) — the ref-count contract of internal.rs:186-209, :450-525 under stress."""
import gc

import numpy as np
import pytest

from trace_gen import F32, I32, U32, same_bits
from vkjit_b200.ir import Bop

pytestmark = pytest.mark.gpu


class Side:
    """Mirrors one handle table on one Ir: handle -> var id (None once dropped)."""

    def __init__(self, ir):
        self.ir, self.h = ir, []

    def add(self, vid):
        self.h.append(vid)
        return len(self.h) - 1

    def drop(self, k):
        if self.h[k] is not None:
            self.ir.dec_ref_count(self.h[k])
            self.h[k] = None


@pytest.mark.parametrize("seed", range(8))
def test_random_lifetimes(cuda_backend, cir, oir, seed):
    gc.collect()
    cuda_backend.sync()
    base = cuda_backend.stats()["pool_bytes_live"]
    rng = np.random.default_rng(seed)
    n = int(rng.choice([5, 257, 4099]))
    sides = [Side(cir), Side(oir)]
    tys = []   # type per handle (same on both sides)

    def new_leaf():
        kind = rng.integers(0, 3)
        ty = int(rng.choice([U32, I32, F32]))
        if kind == 0:
            data = rng.integers(0, 1 << 31, n)
            for s in sides:
                s.add({U32: s.ir.array_u32, I32: s.ir.array_i32, F32: s.ir.array_f32}[ty](data))
        elif kind == 1:
            for s in sides:
                s.add(s.ir.arange(ty, n))
        else:
            v = int(rng.integers(1, 100))
            for s in sides:
                s.add({U32: s.ir.const_u32, I32: s.ir.const_i32, F32: s.ir.const_f32}[ty](v))
        tys.append(ty)

    for _ in range(4):
        new_leaf()
    for step in range(60):
        live = [k for k in range(len(tys)) if sides[0].h[k] is not None]
        act = rng.integers(0, 10)
        if act <= 4 and len(live) >= 2:                      # binary op with autocast
            a, b = (int(x) for x in rng.choice(live, 2))
            op = int(rng.choice([Bop.Add, Bop.Sub, Bop.Mul]))
            for s in sides:
                s.add(s.ir.bop(op, s.h[a], s.h[b]))
            tys.append(max(tys[a], tys[b]))
        elif act == 5:
            new_leaf()
        elif act == 6 and len(live) > 3:                     # drop a handle that others may depend on
            k = int(rng.choice(live))
            for s in sides:
                s.drop(k)
        elif act == 7 and live:                              # clone + drop (inc/dec)
            k = int(rng.choice(live))
            for s in sides:
                s.ir.inc_ref_count(s.h[k])
                s.ir.dec_ref_count(s.h[k])
        elif live:                                           # evaluate up to 3 live handles that have a size
            ks = [int(x) for x in rng.choice(live, min(3, len(live)), replace=False)]
            ok = []
            for k in ks:
                try:
                    sides[1].ir.eval([sides[1].h[k]])        # oracle first: skips size-less consts
                    ok.append(k)
                except Exception:
                    pass
            if ok:
                sides[0].ir.eval([sides[0].h[k] for k in ok])
                for k in ok:
                    a = sides[0].ir.as_slice(sides[0].h[k], tys[k])
                    b = sides[1].ir.as_slice(sides[1].h[k], tys[k])
                    assert same_bits(a, b, tys[k] == F32), (seed, step, k)
    # final: everything still alive evaluates to the same values, then everything is released
    for k in range(len(tys)):
        if sides[0].h[k] is None:
            continue
        try:
            sides[1].ir.eval([sides[1].h[k]])
        except Exception:
            continue
        sides[0].ir.eval([sides[0].h[k]])
        assert same_bits(sides[0].ir.as_slice(sides[0].h[k], tys[k]), sides[1].ir.as_slice(sides[1].h[k], tys[k]), tys[k] == F32)
    for k in range(len(tys)):
        sides[0].drop(k)
    cuda_backend.sync()
    assert cir.num_arrays() == 0
    assert cuda_backend.stats()["pool_bytes_live"] == base
