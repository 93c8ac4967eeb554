"""A Var-like wrapper over any `Ir` (oracle or CUDA) with the coercion rules of the vkjit Python
front-end, so front-end style programs (monte_carlo.build) can be replayed on the CPU oracle."""
import numpy as np

from vkjit_b200.ir import Bop, Uop


class IrModule:
    def __init__(self, ir):
        self.ir = ir

    def wrap(self, x):
        if isinstance(x, V):
            return x
        if isinstance(x, bool):
            return V(self, self.ir.const_u32(int(x)))
        if isinstance(x, (int, np.integer)):
            return V(self, self.ir.const_u32(int(x)) if x >= 0 else self.ir.const_i32(int(x)))
        return V(self, self.ir.const_f32(float(x)))

    def arange(self, ty, n):
        return V(self, self.ir.arange(ty, n))

    def select(self, c, a, b):
        return V(self, self.ir.select(self.wrap(c).id, self.wrap(a).id, self.wrap(b).id))

    def maximum(self, a, b):
        return self.wrap(a)._b(Bop.Max, b)

    def minimum(self, a, b):
        return self.wrap(a)._b(Bop.Min, b)

    def _u(self, k, x):
        return V(self, self.ir.uop(k, self.wrap(x).id))

    def sqrt(self, x): return self._u(Uop.Sqrt, x)
    def exp(self, x): return self._u(Uop.Exp, x)
    def log(self, x): return self._u(Uop.Log, x)
    def sin(self, x): return self._u(Uop.Sin, x)
    def cos(self, x): return self._u(Uop.Cos, x)


class V:
    def __init__(self, m, id):
        self.m, self.id = m, id

    def _b(self, k, o, swap=False):
        o = self.m.wrap(o)
        a, b = (o.id, self.id) if swap else (self.id, o.id)
        return V(self.m, self.m.ir.bop(k, a, b))

    def cast(self, ty): return V(self.m, self.m.ir.cast(self.id, ty))
    def __add__(self, o): return self._b(Bop.Add, o)
    def __sub__(self, o): return self._b(Bop.Sub, o)
    def __mul__(self, o): return self._b(Bop.Mul, o)
    def __truediv__(self, o): return self._b(Bop.Div, o)
    def __radd__(self, o): return self._b(Bop.Add, o, True)
    def __rsub__(self, o): return self._b(Bop.Sub, o, True)
    def __rmul__(self, o): return self._b(Bop.Mul, o, True)
    def __xor__(self, o): return self._b(Bop.Xor, o)
    def __and__(self, o): return self._b(Bop.And, o)
    def __or__(self, o): return self._b(Bop.Or, o)
    def __rshift__(self, o): return self._b(Bop.Shr, o)
    def __lshift__(self, o): return self._b(Bop.Shl, o)
    def __gt__(self, o): return self._b(Bop.Gt, o)
    def __lt__(self, o): return self._b(Bop.Lt, o)
    def __ge__(self, o): return self._b(Bop.Geq, o)
    def __le__(self, o): return self._b(Bop.Leq, o)
