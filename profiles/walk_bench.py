"""Host cost of the per-eval trace walk (build_program: the cache-hit critical path) on the 364-node Monte-Carlo
trace; needs no GPU.  `fresh` = one walk right after the Python front-end built the trace (vkjit_debug_walk_ns with
reps = 0), `hot` = mean of 3000 walks of the same trace.
    python profiles/walk_bench.py"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monte_carlo  # noqa: E402
from vkjit_b200 import vkjit  # noqa: E402

ir = vkjit._global_ir()
ns, nodes = C.c_uint64(), C.c_uint32()
fresh, hot = [], []
for i in range(420):
    y = monte_carlo.build(vkjit, 1 << 26, 5)
    ids = (C.c_uint32 * 1)(y.id())
    ir.api.call("debug_walk_ns", ir._h, ids, 1, 0, C.byref(ns), C.byref(nodes))
    fresh.append(ns.value)
    if i % 40 == 0:
        ir.api.call("debug_walk_ns", ir._h, ids, 1, 3000, C.byref(ns), C.byref(nodes))
        hot.append(ns.value)
fresh = sorted(fresh[20:])
hot.sort()
print({"nodes": nodes.value, "fresh_walk_ns": {"min": fresh[0], "median": fresh[len(fresh) // 2], "p90": fresh[int(len(fresh) * .9)]},
       "hot_walk_ns": {"min": hot[0], "median": hot[len(hot) // 2]}})
