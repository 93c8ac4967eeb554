"""Timing of prefix sum / compress at 2^28 u32 (VKJIT_SCAN_DIAG=nolookback: pipeline-only diagnostic)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
ir = Ir()
fb = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
n = 1 << 28
lanes = ir.arange(T.U32, n)
vals = hash_trace(ir, lanes, 3)
mask = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), ir.const_u32(1)), ir.const_u32(0))
ir.eval([vals, mask])
def timed(fn):
    ts = []
    for i in range(9):
        with torch.cuda.stream(stream):
            fb.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); vk.sync()
        if i >= 2: ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)
def scan():
    ir.dec_ref_count(ir.prefix_sum(vals, True))
def comp():
    r, k = ir.compress_values(vals, mask); ir.dec_ref_count(r)
def compi():
    r, k = ir.compress(mask); ir.dec_ref_count(r)
ms, mc, mi = timed(scan), timed(comp), timed(compi)
print(f"diag={os.environ.get('VKJIT_SCAN_DIAG','-')}  prefix_sum {ms:.4f} ms ({8*n/ms/1e6:.0f} GB/s, {8*n/ms/1e6/6450:.3f})  compress {mc:.4f} ms ({10*n/mc/1e6:.0f} GB/s, {10*n/mc/1e6/6450:.3f})  compress_index {mi:.4f} ms ({6*n/mi/1e6:.0f} GB/s of 6 B/lane, {6*n/mi/1e6/6450:.3f})")
