"""Timing of the fused trace -> scan / compress kernels (scan_fused.cuh) at 2^28 lanes.
VKJIT_SCAN_T=512|1024 selects two 512-thread CTAs per SM or one 1024-thread CTA."""
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
ir.eval([vals])
c = ir.const_u32
def timed(fn):
    ts = []
    for i in range(8):
        with torch.cuda.stream(stream):
            fb.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); vk.sync()
        if i >= 2: ts.append(a.elapsed_time(b))
    return sum(ts) / len(ts)
def hash_mask():   # values streamed, mask = hash of the lane index
    mk = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), c(1)), c(0))
    r, k = ir.compress_values(vals, mk); ir.dec_ref_count(r); ir.dec_ref_count(mk)
def thresh():      # values streamed, mask = values > t
    mk = ir.gt(vals, c(0x80000000))
    r, k = ir.compress_values(vals, mk); ir.dec_ref_count(r); ir.dec_ref_count(mk)
def thresh_idx():  # indices of values > t
    mk = ir.gt(vals, c(0x80000000))
    r, k = ir.compress(mk); ir.dec_ref_count(r); ir.dec_ref_count(mk)
def scan_trace():  # prefix sum of hash(lane): nothing read
    h = hash_trace(ir, lanes, 3)
    r = ir.prefix_sum(h, True); ir.dec_ref_count(r); ir.dec_ref_count(h)
def scan_x2():     # prefix sum of (values >> 3): one streamed array
    h = ir.shr(vals, c(3))
    r = ir.prefix_sum(h, True); ir.dec_ref_count(r); ir.dec_ref_count(h)
mask_arr = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), c(1)), c(0))
ir.eval([mask_arr])
def two_streams():  # mask array AND values array both streamed through the ring (12288-lane tiles)
    mk = ir.bop(Bop.And, mask_arr, ir.const_bool(True))
    r, k = ir.compress_values(vals, mk); ir.dec_ref_count(r); ir.dec_ref_count(mk)
out = [f"T={os.environ.get('VKJIT_SCAN_T', '1024')}"]
for name, fn, bpl in (("hash_mask", hash_mask, 6), ("thresh", thresh, 6), ("thresh_idx", thresh_idx, 6), ("scan_trace", scan_trace, 4), ("scan_stream", scan_x2, 8), ("two_streams", two_streams, 10)):
    ms = timed(fn)
    out.append(f"{name} {ms:.4f} ms ({bpl*n/ms/1e6/6450:.3f} of peak at {bpl} B/lane)")
print("  ".join(out))
