"""Small run of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk  # noqa: E402
from vkjit_b200.ir import Bop, Ir, Red, VarType as T  # noqa: E402

vk.init(0)
ir = Ir()
n = 3 * 16384 + 777            # several look-back tiles + a ragged tail
rng = np.random.default_rng(0)
x = ir.array_f32(rng.random(n, dtype=np.float32))
u = ir.array_u32(rng.integers(0, 1 << 32, n, dtype=np.uint64).astype(np.uint32))
z = ir.add(ir.mul(x, x), ir.const_f32(0.5))
ir.eval([z])
for r in (Red.Sum, Red.Min, Red.Max):
    ir.reduce(r, z); ir.reduce(r, u)
    ir.reduce(r, ir.mul(u, ir.const_u32(3)))      # fused trace -> reduce
s = ir.prefix_sum(u, True)
m = ir.neq(ir.bop(Bop.And, u, ir.const_u32(1)), ir.const_u32(0))
c, k = ir.compress_values(u, m)
i, k2 = ir.compress(m)
idx = ir.bop(Bop.And, u, ir.const_u32(1023))
bins = ir.array_u32(np.zeros(1024, np.uint32))
sa = ir.scatter_add(ir.gather(u, idx), bins, idx)
ir.eval([sa])
ref = np.cumsum(ir.as_slice(u, T.U32).astype(np.uint64)).astype(np.uint32)
assert ir.as_slice(s, T.U32)[-1] == ref[-2] and k == k2
# fused trace -> scan / compress kernels (scan_fused.cuh): 0, 1 and 2 streamed arrays, ragged tail, gather in the trace
lanes = ir.arange(T.U32, n)
hsh = ir.mul(lanes, ir.const_u32(2654435761))
fs0 = ir.prefix_sum(hsh, False)                                             # nothing streamed
fm = ir.gt(u, ir.const_u32(1 << 31))
fc, fk = ir.compress_values(u, fm)                                          # one streamed array, values re-read from the ring
fx, fk2 = ir.compress(ir.bop(Bop.And, fm, ir.gt(x, ir.const_f32(0.5))))     # two streamed arrays (12288-lane tiles)
fg = ir.prefix_sum(ir.gather(u, idx), True)                                 # streamed index array + gather pointer
uu = ir.as_slice(u, T.U32)
assert fk == int((uu > (1 << 31)).sum()) and np.array_equal(ir.as_slice(fc, T.U32), uu[uu > (1 << 31)])
assert np.array_equal(ir.as_slice(fs0, T.U32), np.cumsum((np.arange(n, dtype=np.uint64) * 2654435761) & 0xFFFFFFFF).astype(np.uint32))
assert ir.as_slice(fg, T.U32)[1] == uu[int(uu[0]) & 1023]
# shared-memory-privatised scatter_add variant (launches of >= 2^22 lanes)
big = (1 << 22) + 3
bi = ir.bop(Bop.And, ir.mul(ir.arange(T.U32, big), ir.const_u32(2654435761)), ir.const_u32(0xFFFF))
bb = ir.array_u32(np.zeros(1 << 16, np.uint32))
ir.eval([ir.scatter_add(ir.const_u32(1), bb, bi)])
assert int(ir.as_slice(bb, T.U32).astype(np.uint64).sum()) == big
vk.sync()
print("sanitize workload ok", k)
