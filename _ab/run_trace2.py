import os, sys
sys.path.insert(0, '/root/repo')
import numpy as np
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Ir, VarType as T
vk.init(0)
ir = Ir()
n = 1 << 28
vals = hash_trace(ir, ir.arange(T.U32, n), 3)
ir.eval([vals])
ref = None
for i in range(3):
    r = ir.prefix_sum(vals, True)
    if i == 0:
        got = ir.as_slice(r, T.U32); v = ir.as_slice(vals, T.U32)
        exp = np.cumsum(v.astype(np.uint64)).astype(np.uint32)
        print("correct:", bool(got[0] == 0 and np.array_equal(got[1:], exp[:-1])))
    ir.dec_ref_count(r)
vk.sync()
t = np.fromfile(os.environ['VKJIT_SCAN_TRACE'], dtype=np.uint64).reshape(-1, 8).astype(np.int64)[:-1]
t0 = t[:, 0].min()
start, data, refill, lbdone, done, cta, lbstart = [t[:, i] - t0 for i in (0, 1, 2, 3, 4, 5, 6)]
print("tiles", len(t), "total us", done.max() / 1e3)
for name, d in (("tma_wait", data - start), ("scanA", refill - data), ("resolve", lbdone - lbstart), ("output", done - lbdone), ("start->done", done - start)):
    print(f"{name:11s} mean {d.mean():8.0f} ns  p50 {np.percentile(d,50):8.0f}  p90 {np.percentile(d,90):8.0f}  p99 {np.percentile(d,99):8.0f}")
c0 = t[cta == 0]
per = np.diff(np.sort(c0[:, 0]))
print("CTA0 iteration period ns: mean", per.mean(), "p50", np.median(per))
print("CTA0 timeline (us, relative to tile start of g=30): cols start,data,refill,lbstart,lbdone,done")
base = t[148 * 30, 0]
for g in range(30, 35):
    r = t[148 * g]
    print(g, [round((r[i] - base) / 1e3, 2) for i in (0, 1, 2, 6, 3, 4)])
