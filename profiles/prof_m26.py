"""Launches the Monte-Carlo mega-trace kernel (BASELINE configs[4], 2^26 lanes) three times for one ncu capture:
    ncu --clock-control none -k regex:vkjit_trace -s 1 -c 1 --metrics smsp__inst_executed.sum,... python profiles/prof_m26.py
Never a bench."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import monte_carlo  # noqa: E402
import vkjit_b200 as vk  # noqa: E402
from vkjit_b200 import vkjit  # noqa: E402

vk.init(0)
for _ in range(3):
    y = monte_carlo.build(vkjit, 1 << 26, 5)
    vkjit.eval([y])
    vk.sync()
print("prof_m26 done")
