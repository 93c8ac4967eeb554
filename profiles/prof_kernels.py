"""Launches every hot kernel a few times at the BASELINE sizes so that one ncu capture sees them
all (run under `ncu --set full -k regex:... python profiles/prof_kernels.py`).  Never a bench."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk  # noqa: E402
from bench import hash_trace, uniform_trace  # noqa: E402
from vkjit_b200.ir import Bop, Ir, Red, VarType as T  # noqa: E402

which = set(sys.argv[1:]) or {"reduce", "trace", "scan", "compress", "hist"}
vk.init(0)
ir = Ir()
n = 1 << int(os.environ.get("PROF_LOG2N", "28"))
lanes = ir.arange(T.U32, n)
x = uniform_trace(ir, lanes, 1)
y = uniform_trace(ir, lanes, 2)
ir.eval([x, y])
vk.sync()
reps = 2
if "reduce" in which:
    for _ in range(reps):
        ir.dec_ref_count(ir.reduce(Red.Sum, x))
        ir.dec_ref_count(ir.reduce(Red.Max, x))
if "trace" in which:
    half = ir.const_f32(0.5)
    for _ in range(reps):
        z = ir.add(ir.mul(x, y), half)
        ir.eval([z])
        ir.dec_ref_count(z)
if "scan" in which or "compress" in which:
    vals = hash_trace(ir, lanes, 3)
    mask = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), ir.const_u32(1)), ir.const_u32(0))
    ir.eval([vals, mask])
    for _ in range(reps):
        if "scan" in which:
            ir.dec_ref_count(ir.prefix_sum(vals, True))
        if "compress" in which:
            r, k = ir.compress_values(vals, mask)
            ir.dec_ref_count(r)
if "hist" in which:
    m = 1 << 26
    l26 = ir.arange(T.U32, m)
    idx = ir.bop(Bop.And, hash_trace(ir, l26, 5), ir.const_u32(0xFFFF))
    table = hash_trace(ir, ir.arange(T.U32, 1 << 16), 6)
    ir.eval([idx]); ir.eval([table])
    bins = ir.array_u32(np.zeros(1 << 16, np.uint32))
    for _ in range(reps):
        s = ir.scatter_add(ir.gather(table, idx), bins, idx)
        ir.eval([s])
        ir.dec_ref_count(s)
vk.sync()
print("prof_kernels done", sorted(which))
