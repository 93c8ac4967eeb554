"""H26 A/B: privatised scatter_add per CTA (192 / 224 KB of bins in shared memory, the rest through L2 RED) against the
2-CTA-cluster variant (all 2^16 bins in distributed shared memory, 128 KB per CTA).  VKJIT_SADD_CLUSTER=0/1 per process."""
import json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
fb = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
ir = Ir(); c = ir.const_u32
m = 1 << 26
lanes = ir.arange(T.U32, m)
idx = ir.bop(Bop.And, hash_trace(ir, lanes, 0xB2000031), c(0xFFFF)); ir.eval([idx])
hh = hash_trace(ir, lanes, 0xB2000031)
idx_s = ir.bop(Bop.Min, ir.bop(Bop.And, hh, c(0xFFFF)), ir.shr(hh, c(16))); ir.eval([idx_s])
table = hash_trace(ir, ir.arange(T.U32, 1 << 16), 0xB2000032); ir.eval([table])
bins = ir.array_u32(np.zeros(1 << 16, np.uint32))
one = c(1)
def timed(fn):
    ts = []
    for i in range(8):
        with torch.cuda.stream(stream):
            fb.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); vk.sync()
        if i >= 2: ts.append(a.elapsed_time(b))
    return round(sorted(ts)[len(ts) // 2], 4)
def gs(ix):
    s = ir.scatter_add(ir.gather(table, ix), bins, ix); ir.eval([s]); ir.dec_ref_count(s)
def cnt(ix):
    s = ir.scatter_add(one, bins, ix); ir.eval([s]); ir.dec_ref_count(s)
out = {"VKJIT_SADD_CLUSTER": os.environ.get("VKJIT_SADD_CLUSTER", "default"), "VKJIT_SADD_PASSES": os.environ.get("VKJIT_SADD_PASSES", "default"), "VKJIT_PASS_KB": os.environ.get("VKJIT_PASS_KB", "default"),
       "gather_scatter_add_ms": timed(lambda: gs(idx)), "count_ms": timed(lambda: cnt(idx)),
       "skewed_gather_scatter_add_ms": timed(lambda: gs(idx_s))}
total = int(ir.as_slice(bins, T.U32).astype(np.uint64).sum() % (1 << 32))
out["bins_checksum"] = total
print(json.dumps(out))
