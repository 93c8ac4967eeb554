"""Where a tile period of the lagged fused trace -> compress kernel goes (scan_fused.cuh, VK_TRACE stamps).
    VKJIT_FSCAN_TRACE=/tmp/fscan.bin python profiles/fscan_timeline.py [thresh|hash_mask|thresh_idx]
Stamps are %globaltimer values written by thread 0 (warp 0: also the serial section) and thread 64 (a plain worker
warp) into the spare words of each tile's status line; medians over all tiles of the steady state, in us."""
import os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
which = sys.argv[1] if len(sys.argv) > 1 else "thresh"
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
        mk = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 4), c(1)), c(0))
        r, k = ir.compress_values(vals, mk)
    elif which == "thresh_idx":
        mk = ir.gt(vals, c(0x80000000))
        r, k = ir.compress(mk)
    else:
        mk = ir.gt(vals, c(0x80000000))
        r, k = ir.compress_values(vals, mk)
    ir.dec_ref_count(r); ir.dec_ref_count(mk)
for _ in range(3):
    run()
vk.sync()
w = np.fromfile(tf, dtype=np.uint64).reshape(-1, 16)[1:]   # line 0 is the scratch header
tiles = len(w)
cta = w[:, 10].astype(np.int64)
grid = int(cta.max()) + 1
st = w.astype(np.int64)
per = tiles // grid
t = np.arange(tiles)
nxt = t + grid
ok = (nxt < tiles) & (t >= 4 * grid) & (nxt < tiles - 4 * grid)    # steady state, and the tile has a successor in its CTA
cur, nx = st[t[ok]], st[nxt[ok]]
def med(x): return float(np.median(x)) / 1e3
rows = [
    ("period (start of tile k -> start of tile k+1 in the same CTA)", nx[:, 1] - cur[:, 1]),
    ("wait for the tile's data (TMA)", cur[:, 2] - cur[:, 1]),
    ("evaluate + local scan (thread 0)", cur[:, 3] - cur[:, 2]),
    ("thread 0 at barrier 1", cur[:, 4] - cur[:, 3]),
    ("worker warp: iteration start -> reaches barrier 1", cur[:, 11] - cur[:, 1]),
    ("worker warp waits at barrier 1", cur[:, 4] - cur[:, 11]),
    ("warp 0: scan of the totals + publish", cur[:, 5] - cur[:, 4]),
    ("warp 0: resolve the previous tile's prefix (tile k-1 of this CTA, same iteration)", nx[:, 5] * 0 + (cur[:, 6] - nx[:, 5])),
    ("worker warp: barrier 1 passed -> barrier 2 passed (serial section seen by a worker)", cur[:, 7] - nx[:, 4]),
    ("worker warp: output of the tile", cur[:, 8] - cur[:, 7]),
    ("worker warp: slot-release barrier", cur[:, 9] - cur[:, 8]),
]
print(f"case {which}: {tiles} tiles, grid {grid}, {per} tiles per CTA, kernel span {(st[:, 1:10][st[:, 1:10] > 0].max() - st[:, 1][st[:, 1] > 0].min()) / 1e3:.1f} us")
for name, x in rows:
    print(f"  {med(x):7.2f} us  {name}")
