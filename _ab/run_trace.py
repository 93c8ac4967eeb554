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
for i in range(3):
    ir.dec_ref_count(ir.prefix_sum(vals, True))
vk.sync()
t = np.fromfile(os.environ['VKJIT_SCAN_TRACE'], dtype=np.uint64).reshape(-1, 8).astype(np.int64)
tiles = t.shape[0]
t0 = t[:, 0].min()
start, data, phA, lb, done, cta = (t[:, i] - t0 for i in range(5)), None, None, None, None, None
start, data, phA, lb, done = [t[:, i] - t0 for i in range(5)]
cta = t[:, 5]
print("tiles", tiles, "total us", (done.max()) / 1e3)
d_wait = data - start; d_A = phA - data; d_lb = lb - phA; d_out = done - lb
for name, d in (("tma_wait", d_wait), ("phaseA", d_A), ("lookback", d_lb), ("output", d_out), ("tile", done - start)):
    print(f"{name:9s} mean {d.mean():8.0f} ns  p50 {np.percentile(d,50):8.0f}  p90 {np.percentile(d,90):8.0f}  p99 {np.percentile(d,99):8.0f}  max {d.max()}")
# generation view
G = 148
gens = tiles // G
pa = phA[:gens * G].reshape(gens, G); lbd = lb[:gens * G].reshape(gens, G)
print("per generation: spread of aggregate-publish times (max-min) mean", (pa.max(1) - pa.min(1)).mean(), "ns; p50 of (lookback_done - max publish in gen)", np.median(lbd.max(1) - pa.max(1)))
# lookback duration vs position in generation
pos = np.arange(tiles) % G
for lo in (0, 1, 8, 32, 64, 100, 140):
    sel = (pos >= lo) & (pos < lo + 8)
    print(f"pos {lo:3d}-{lo+7:3d}: lookback mean {d_lb[sel].mean():7.0f} ns, publish offset within gen mean {(phA[sel] - np.repeat(pa.min(1), G)[:tiles][sel] if False else 0)}")
gi = 30
print("gen", gi, "publish (us rel):", np.round((pa[gi] - pa[gi].min()) / 1e3, 2)[:20], "...", np.round((pa[gi] - pa[gi].min()) / 1e3, 2)[-8:])
print("gen", gi, "lb done (us rel):", np.round((lbd[gi] - pa[gi].min()) / 1e3, 2)[:20], "...", np.round((lbd[gi] - pa[gi].min()) / 1e3, 2)[-8:])
print("gen period us:", np.round(np.diff(pa.min(1))[:40] / 1e3, 2))
