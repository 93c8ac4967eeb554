"""One fused trace -> compress launch at 2^28 lanes for ncu (kernel name vkjit_trace, second launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Ir, VarType as T
vk.init(0)
ir = Ir()
n = 1 << 28
vals = hash_trace(ir, ir.arange(T.U32, n), 3)
ir.eval([vals])                                   # vkjit_trace launch 0
mk = ir.gt(vals, ir.const_u32(0x80000000))
r, k = ir.compress_values(vals, mk)               # vkjit_trace launch 1: the fused kernel
print("selected", k)
