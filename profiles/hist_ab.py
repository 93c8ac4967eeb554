"""H26 timing (gather + scatter-add and counting histogram, 2^26 indices into 2^16 bins); VKJIT_PRIV_KB sets the
shared-memory budget of the privatised bins, VKJIT_NO_PRIVATIZE=1 selects the plain L2 RED path."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0); ir = Ir(); stream = torch.cuda.ExternalStream(vk.stream_ptr())
m = 1 << 26
idx = ir.bop(Bop.And, hash_trace(ir, ir.arange(T.U32, m), 5), ir.const_u32(0xFFFF))
table = hash_trace(ir, ir.arange(T.U32, 1 << 16), 6)
ir.eval([idx]); ir.eval([table])
bins = ir.array_u32(np.zeros(1 << 16, np.uint32))
out = []
for variant in ("gather+scatter_add", "count"):
    ts = []
    for i in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        s = ir.scatter_add(ir.gather(table, idx) if variant[0] == "g" else ir.const_u32(1), bins, idx); ir.eval([s])
        b.record(stream); vk.sync(); ir.dec_ref_count(s)
        if i >= 2: ts.append(a.elapsed_time(b))
    out.append(f"{variant}: {sum(ts)/len(ts):.4f} ms {m/(sum(ts)/len(ts))/1e6:.0f} Gelem/s")
print(f"PRIV_KB={os.environ.get('VKJIT_PRIV_KB','192')} NO_PRIV={os.environ.get('VKJIT_NO_PRIVATIZE','0')}  " + "  |  ".join(out))
