"""Random construct / clone / drop sequences WITHOUT a device: the product Ir (flat var table, free list, a cascade
that only stacks dying vars) against an exact reference-count model — internal.rs:186-209 (a new var starts at 1 and
adds 1 to every dependency) and :450-469 (at 0 a var releases its dependencies), with the one documented
difference that EVERY dependency edge is released: the reference's visitor keeps a `discovered` set per cascade
(iterators.rs:63-88), so a var reached twice in one cascade (listed twice by a node, or through a diamond) is
decremented once and leaks.  The programs avoid implicit casts, which the reference leaks as well."""
import numpy as np
import pytest

from trace_gen import BOOL, F32, I32, U32
from vkjit_b200.ir import Bop, Ir, Uop


class Model:
    """Edge-exact reference counting over model ids (independent of the product's var ids)."""

    def __init__(self):
        self.rc, self.deps, self.sized = {}, {}, {}
        self.next = 0

    def new(self, deps=(), lanes=False):
        i = self.next
        self.next += 1
        self.rc[i], self.deps[i] = 1, list(deps)
        self.sized[i] = lanes or any(self.sized[d] for d in deps)   # reaches an Arange: a schedule needs a kernel size
        for d in deps:
            self.rc[d] += 1
        return i

    def inc(self, i):
        self.rc[i] += 1

    def dec(self, i):
        stack = [i]
        while stack:
            v = stack.pop()
            assert self.rc[v] > 0
            self.rc[v] -= 1
            if self.rc[v] == 0:
                stack.extend(self.deps[v])
                self.deps[v] = []

    def evaluate(self, roots):
        """Ir::eval's bookkeeping (internal.rs:482-525): +1 per scheduled root, every root gives up its dependency
        edges and becomes a leaf (a Binding), the schedule's references are dropped again."""
        sched = []
        for r in roots:
            if r not in sched:                      # duplicate roots collapse
                sched.append(r)
                self.rc[r] += 1
        released = []
        for r in sched:
            released += self.deps[r]
            self.deps[r] = []
        for d in released:
            self.dec(d)
        for r in sched:
            self.dec(r)

    def live(self):
        return sum(1 for c in self.rc.values() if c > 0)


def live_vars(ir):
    text = ir.repr()
    return text.count("Var {") - text.count("op: Free")


@pytest.mark.parametrize("seed", range(16))
def test_random_handle_lifetimes_match_the_model(seed):
    rng = np.random.default_rng(1000 + seed)
    ir, model = Ir(), Model()
    hp, hm, tys = [], [], []          # per handle: product var id, model id, type (None once dropped)

    def add(pid, mid, ty):
        hp.append(pid); hm.append(mid); tys.append(ty)

    def leaf():
        ty = int(rng.choice([U32, I32, F32]))
        if rng.random() < 0.5:
            add(ir.arange(ty, 16), model.new(lanes=True), ty)
        else:
            v = int(rng.integers(1, 100))
            add({U32: ir.const_u32, I32: ir.const_i32, F32: ir.const_f32}[ty](v), model.new(), ty)

    def check(tag):
        for k in range(len(tys)):
            if hp[k] is not None:
                assert ir.ref_count(hp[k]) == model.rc[hm[k]], (seed, tag, k)
        assert live_vars(ir) == model.live(), (seed, tag)

    for _ in range(3):
        leaf()
    for step in range(200):
        live = [k for k in range(len(tys)) if hp[k] is not None]
        num = [k for k in live if tys[k] in (U32, I32, F32)]

        def same_type(a):
            return [k for k in num if tys[k] == tys[a]]

        act = int(rng.integers(0, 12))
        if act <= 3 and num:                                         # arithmetic between equal types; a op a allowed
            a = int(rng.choice(num))
            b = int(rng.choice(same_type(a)))
            op = int(rng.choice([Bop.Add, Bop.Sub, Bop.Mul, Bop.Min, Bop.Max]))
            add(ir.bop(op, hp[a], hp[b]), model.new([hm[a], hm[b]]), tys[a])
        elif act == 4 and num:                                       # comparison -> Bool
            a = int(rng.choice(num))
            b = int(rng.choice(same_type(a)))
            add(ir.bop(Bop.Lt, hp[a], hp[b]), model.new([hm[a], hm[b]]), BOOL)
        elif act == 5 and num:                                       # select
            bools = [k for k in live if tys[k] == BOOL]
            if bools:
                a = int(rng.choice(num))
                b, c = int(rng.choice(same_type(a))), int(rng.choice(bools))
                add(ir.select(hp[c], hp[a], hp[b]), model.new([hm[c], hm[a], hm[b]]), tys[a])
        elif act == 6 and num:                                       # cast: same type returns the operand itself, no new count
            a, ty = int(rng.choice(num)), int(rng.choice([U32, I32, F32]))
            out = ir.cast(hp[a], ty)
            if ty == tys[a]:
                assert out == hp[a]
            else:
                add(out, model.new([hm[a]]), ty)
        elif act == 7 and num:                                       # unary
            a = int(rng.choice(num))
            add(ir.uop(Uop.Neg, hp[a]), model.new([hm[a]]), tys[a])
        elif act == 8 and num:                                       # struct round trip: init -> setattr -> getattr (diamonds)
            a = int(rng.choice(num))
            b = int(rng.choice(same_type(a)))
            st, mst = ir.struct_init([hp[a], hp[b]]), model.new([hm[a], hm[b]])
            st2, mst2 = ir.setattr(st, hp[b], 0), model.new([mst, hm[b]])
            g, mg = ir.getattr(st2, 1), model.new([mst2])
            for v in (st, st2):
                ir.dec_ref_count(v)
            for v in (mst, mst2):
                model.dec(v)
            add(g, mg, tys[b])
        elif act == 9 and live:                                      # clone, sometimes kept as a handle of its own
            k = int(rng.choice(live))
            ir.inc_ref_count(hp[k]); model.inc(hm[k])
            if rng.random() < 0.5:
                add(hp[k], hm[k], tys[k])
            else:
                ir.dec_ref_count(hp[k]); model.dec(hm[k])
        elif act == 10 and num:                                      # eval bookkeeping of up to 3 sized roots (no device)
            import ctypes as C
            ks = [int(x) for x in rng.choice(num, min(3, len(num)), replace=False)]
            ks = [k for k in ks if model.sized[hm[k]]]
            if ks:
                ids = (C.c_uint32 * len(ks))(*[hp[k] for k in ks])
                ir.api.call("debug_eval_bookkeeping", ir._h, ids, len(ks))
                model.evaluate([hm[k] for k in ks])
                for k in ks:
                    assert ir.is_buffer(hp[k])
        elif act >= 10 and len(live) > 2:                            # drop a handle others may depend on
            k = int(rng.choice(live))
            ir.dec_ref_count(hp[k]); model.dec(hm[k])
            hp[k] = None
        else:
            leaf()
        if step % 5 == 0:
            check(step)
    check("end")
    for k in range(len(tys)):
        if hp[k] is not None:
            ir.dec_ref_count(hp[k]); model.dec(hm[k])
            hp[k] = None
    assert live_vars(ir) == 0 and model.live() == 0   # every var was released (the free list holds all slots)
    with pytest.raises(Exception):
        ir.dec_ref_count(0)                           # and a released var cannot be released again
    ir.close()
