"""Per-phase host time of a cache-hit vkjit_eval ($VKJIT_EVAL_TRACE=1: walk / lookup / alloc / launch / commit on stderr)
for the 364-node Monte-Carlo trace (2^26 lanes) and for a 3-op trace (n = 1024).
    VKJIT_EVAL_TRACE=1 python profiles/eval_phases.py 2> gpurun_out/eval_phases.txt"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monte_carlo  # noqa: E402
import vkjit_b200 as vk  # noqa: E402
from vkjit_b200 import vkjit  # noqa: E402

vk.init(0)
for log2n in (26, 14):
    print(f"--- monte carlo 2^{log2n}", file=sys.stderr, flush=True)
    for i in range(8):
        y = monte_carlo.build(vkjit, 1 << log2n, 5)
        vkjit.eval([y])
        print("    last_eval_ns", vk.stats()["last_eval_ns"], file=sys.stderr, flush=True)
    vk.sync()
print("--- 3-op trace n=1024", file=sys.stderr, flush=True)
a = vkjit.arange(5, 1024)
for i in range(8):
    z = a * 0.5 + 0.5
    vkjit.eval([z])
    print("    last_eval_ns", vk.stats()["last_eval_ns"], file=sys.stderr, flush=True)
vk.sync()
