"""Seeded random trace generator: builds the SAME DAG on any `Ir` (oracle or CUDA).

Used by the parity tests: ints / masks / indices must match bit for bit, f32 + - * / sqrt
and casts too (IEEE, no contraction on either side); NaNs compare equal to NaNs (payloads
differ between x86 SSE and the GPU).
"""
import numpy as np

from vkjit_b200.ir import Bop, Uop, VarType

F32, U32, I32, BOOL = VarType.F32, VarType.U32, VarType.I32, VarType.Bool
NUM = (U32, I32, F32)


def special_f32(rng, n):
    a = rng.standard_normal(n).astype(np.float32) * np.float32(10.0) ** rng.integers(-3, 4, n).astype(np.float32)
    specials = np.array([0.0, -0.0, 1.0, -1.0, 0.5, 3.0, 1e-38, -1e-38, 1e38, 16777216.0, 16777217.0, 4294967296.0,
                         -2147483648.0, 2147483520.0, 0.1, 1e-45], dtype=np.float32)
    k = min(n, len(specials))
    pos = rng.choice(n, k, replace=False)
    a[pos] = specials[:k]
    return a


def special_u32(rng, n):
    a = rng.integers(0, 2 ** 32, n, dtype=np.uint64).astype(np.uint32)
    specials = np.array([0, 1, 2, 31, 32, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF, 0xFFFF, 0x10000], dtype=np.uint32)
    k = min(n, len(specials))
    pos = rng.choice(n, k, replace=False)
    a[pos] = specials[:k]
    return a


class TraceBuilder:
    """Replays a fixed random program; `ir` only receives the resulting calls."""

    def __init__(self, ir, seed, n, n_ops=24, transcendental=False, allow_div=True, arrays=True):
        self.ir, self.n = ir, n
        self.rng = np.random.default_rng(seed)
        self.pool = {U32: [], I32: [], F32: [], BOOL: []}
        self.transcendental = transcendental
        self.allow_div = allow_div
        self.n_ops = n_ops
        self.arrays = arrays  # False: only arange/const leaves (traces that need no device memory)

    def leaf(self, ty):
        r, ir = self.rng, self.ir
        kind = r.integers(0, 3)
        if kind == 0 and not self.arrays:
            kind = 1
        if kind == 0:  # array
            if ty == F32:
                return ir.array_f32(special_f32(r, self.n))
            if ty == U32:
                return ir.array_u32(special_u32(r, self.n))
            return ir.array_i32(special_u32(r, self.n).view(np.int32))
        if kind == 1:
            return ir.arange(ty, self.n)
        if ty == F32:
            return ir.const_f32(float(np.float32(r.standard_normal() * 4)))
        if ty == U32:
            return ir.const_u32(int(r.integers(0, 2 ** 32)))
        return ir.const_i32(int(r.integers(-2 ** 31, 2 ** 31)))

    def pick(self, ty):
        p = self.pool[ty]
        if not p or self.rng.random() < 0.25:
            if ty == BOOL:
                a, b = self.pick(self.rng.choice(NUM)), None
                b = self.pick(self.ir.ty(a))
                v = self.ir.bop(int(self.rng.choice([Bop.Lt, Bop.Gt, Bop.Eq, Bop.Leq, Bop.Geq, Bop.Neq])), a, b)
            else:
                v = self.leaf(ty)
            p.append(v)
            return v
        return p[self.rng.integers(0, len(p))]

    def step(self):
        r, ir = self.rng, self.ir
        c = r.integers(0, 12)
        if c <= 3:  # arithmetic with autocast between numeric types
            a, b = self.pick(int(r.choice(NUM))), self.pick(int(r.choice(NUM)))
            ops = [Bop.Add, Bop.Sub, Bop.Mul]
            rt = max(ir.ty(a), ir.ty(b))
            if self.allow_div and rt == F32:
                ops.append(Bop.Div)
            v = ir.bop(int(r.choice(ops)), a, b)
        elif c == 4:  # integer division by a positive constant (x/0 and INT_MIN/-1 are undefined in SPIR-V)
            ty = int(r.choice([U32, I32]))
            a = self.pick(ty)
            d = ir.const_u32(int(r.integers(1, 1000))) if ty == U32 else ir.const_i32(int(r.integers(1, 1000)))
            v = ir.bop(Bop.Div, a, d)
        elif c == 5:  # comparison -> Bool
            ty = int(r.choice(NUM))
            v = ir.bop(int(r.choice([Bop.Lt, Bop.Gt, Bop.Eq, Bop.Leq, Bop.Geq, Bop.Neq])), self.pick(ty), self.pick(ty))
        elif c == 6:  # cast
            src, dst = int(r.choice(NUM)), int(r.choice(NUM))
            v = ir.cast(self.pick(src), dst)
        elif c == 7:  # select
            ty = int(r.choice(NUM))
            v = ir.select(self.pick(BOOL), self.pick(ty), self.pick(ty))
        elif c == 8:  # bit ops / shifts (extension)
            ty = int(r.choice([U32, I32]))
            v = ir.bop(int(r.choice([Bop.And, Bop.Or, Bop.Xor, Bop.Shl, Bop.Shr])), self.pick(ty), self.pick(ty))
        elif c == 9:  # min / max (extension)
            ty = int(r.choice(NUM))
            v = ir.bop(int(r.choice([Bop.Min, Bop.Max])), self.pick(ty), self.pick(ty))
        elif c == 10:  # unary (extension)
            ty = int(r.choice(NUM))
            ops = [Uop.Neg, Uop.Abs]
            if ty != F32:
                ops.append(Uop.Not)
            else:
                ops.append(Uop.Sqrt)
                if self.transcendental:
                    ops += [Uop.Exp, Uop.Log, Uop.Sin, Uop.Cos]
            v = ir.uop(int(r.choice(ops)), self.pick(ty))
        else:  # bool logic / bitcast
            if r.random() < 0.5:
                v = ir.bop(int(r.choice([Bop.And, Bop.Or, Bop.Xor])), self.pick(BOOL), self.pick(BOOL))
            else:
                src, dst = int(r.choice(NUM)), int(r.choice(NUM))
                v = ir.bitcast(self.pick(src), dst)
        self.pool[ir.ty(v)].append(v)
        return v

    def build(self, n_roots=3):
        # make sure the kernel has a size: at least one lane-sized leaf
        self.pool[U32].append(self.ir.arange(U32, self.n))
        last = [self.step() for _ in range(self.n_ops)]
        roots = []
        for v in reversed(last):
            if v not in roots:
                roots.append(v)
            if len(roots) == n_roots:
                break
        return roots


def same_bits(a: np.ndarray, b: np.ndarray, is_f32: bool) -> bool:
    """Bit-exact, except that any NaN equals any NaN."""
    if a.shape != b.shape:
        return False
    au, bu = a.view(np.uint32), b.view(np.uint32)
    if is_f32:
        an, bn = np.isnan(a.view(np.float32)), np.isnan(b.view(np.float32))
        return bool(np.all((au == bu) | (an & bn)))
    return bool(np.all(au == bu))


def ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Distance in units in the last place between two f32 arrays (NaN==NaN -> 0)."""
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -(ai & 0x7FFFFFFF), ai)
    bi = np.where(bi < 0, -(bi & 0x7FFFFFFF), bi)
    d = np.abs(ai - bi)
    both_nan = np.isnan(a) & np.isnan(b)
    return np.where(both_nan, 0, d)
