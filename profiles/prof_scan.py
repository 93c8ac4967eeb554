"""One prefix sum + one compress_values at 2^28 lanes for ncu (-k regex:scan_kernel)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
ir = Ir()
n = 1 << 28
lanes = ir.arange(T.U32, n)
vals = hash_trace(ir, lanes, 3)
mask = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), ir.const_u32(1)), ir.const_u32(0))
ir.eval([vals, mask])
for _ in range(2):
    ir.dec_ref_count(ir.prefix_sum(vals, True))
    r, k = ir.compress_values(vals, mask); ir.dec_ref_count(r)
vk.sync()
print("ok", k)
