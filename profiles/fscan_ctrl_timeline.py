"""Phases of the control-warp fused compress kernel (scan_fused.cuh: VK_CTRL, VK_TRACE stamps), medians in us.
    VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin python profiles/fscan_ctrl_timeline.py [thresh|thresh_idx|hash_mask]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
which = sys.argv[1] if len(sys.argv) > 1 else "thresh_idx"
tf = os.environ["VKJIT_FSCAN_TRACE"]
vk.init(0)
ir = Ir()
n = 1 << 28
lanes = ir.arange(T.U32, n)
vals = hash_trace(ir, lanes, 3)
ir.eval([vals])
c = ir.const_u32
def run():
    if which == "hash_mask":
        mk = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), c(1)), c(0)); r, k = ir.compress_values(vals, mk)
    elif which == "thresh_idx":
        mk = ir.gt(vals, c(0x80000000)); r, k = ir.compress(mk)
    else:
        mk = ir.gt(vals, c(0x80000000)); r, k = ir.compress_values(vals, mk)
    ir.dec_ref_count(r); ir.dec_ref_count(mk)
for _ in range(3):
    run()
vk.sync()
st = np.fromfile(tf, dtype=np.uint64).reshape(-1, 16)[1:].astype(np.int64)
tiles = len(st)
grid = int(st[:, 15].max()) + 1
t = np.arange(tiles)
ok = (t >= 6 * grid) & (t + grid < tiles - 6 * grid)
cur, nx = st[t[ok]], st[t[ok] + grid]
def med(x): return float(np.median(x)) / 1e3
print(f"case {which}: {tiles} tiles, grid {grid}, span {(st[:, 1:13].max() - st[:, 1][st[:, 1] > 0].min()) / 1e3:.1f} us")
for name, x in (
    ("control period (top of iteration j -> top of j+1)", nx[:, 1] - cur[:, 1]),
    ("control: waits for the workers' totals of tile j", cur[:, 2] - cur[:, 1]),
    ("control: scan totals + publish", cur[:, 3] - cur[:, 2]),
    ("control: waits for the window requested last iteration", cur[:, 4] - cur[:, 3]),
    ("control: look-back walk + hand-over", cur[:, 5] - cur[:, 4]),
    ("control: slot release wait + TMA refill", cur[:, 6] - cur[:, 5]),
    ("worker period (start of tile k -> start of k+1)", nx[:, 7] - cur[:, 7]),
    ("worker: waits for the tile's data", cur[:, 8] - cur[:, 7]),
    ("worker: evaluate + row counts + arrive", cur[:, 9] - cur[:, 8]),
    ("worker: waits for the prefix of the tile it writes", cur[:, 11] - cur[:, 10]),
    ("worker: output", cur[:, 12] - cur[:, 11]),
    ("tile: evaluated (worker) -> totals picked up by control", cur[:, 2] - cur[:, 9]),
    ("tile: published -> resolved", cur[:, 5] * 0 + (st[t[ok] + 0, 5] - cur[:, 3])),
):
    print(f"  {med(x):7.2f} us  {name}")
