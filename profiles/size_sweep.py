"""Reduce / trace kernel time vs size (events, L2 evicted by a read) — finds the fixed per-launch cost."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import uniform_trace
from vkjit_b200.ir import Ir, Red, VarType as T
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
ir = Ir()
fb = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
def flush():
    with torch.cuda.stream(stream):
        fb.sum()
for lg in range(16, 29, 2):
    n = 1 << lg
    x = uniform_trace(ir, ir.arange(T.U32, n), 7); ir.eval([x])
    ts = []
    for i in range(13):
        flush()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); r = ir.reduce(Red.Sum, x); b.record(stream); vk.sync()
        ir.dec_ref_count(r)
        if i >= 3: ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    # empty interval: two events back to back
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush(); a.record(stream); b.record(stream); vk.sync()
    print(f"n=2^{lg} bytes={4*n:>11} reduce_us median={ts[len(ts)//2]:.2f} min={ts[0]:.2f}  empty_interval_us={a.elapsed_time(b)*1e3:.2f}  ideal_us@6.9TB/s={4*n/6.9e6:.2f}")
    ir.dec_ref_count(x)
