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
    ("   of which: until the prefetched status window has landed (cp.async.wait_group)", cur[:, 12] - nx[:, 5]),
    ("   of which: the look-back walk itself", cur[:, 6] - cur[:, 12]),
    ("worker warp: barrier 1 passed -> barrier 2 passed (serial section seen by a worker)", cur[:, 7] - nx[:, 4]),
    ("worker warp: output of the tile", cur[:, 8] - cur[:, 7]),
    ("worker warp: slot-release barrier", cur[:, 9] - cur[:, 8]),
]
print(f"case {which}: {tiles} tiles, grid {grid}, {per} tiles per CTA, kernel span {(st[:, 1:10][st[:, 1:10] > 0].max() - st[:, 1][st[:, 1] > 0].min()) / 1e3:.1f} us")
for name, x in rows:
    print(f"  {med(x):7.2f} us  {name}")
rounds, polls = cur[:, 14], cur[:, 13]
print(f"  look-back rounds per tile: mean {rounds.mean():.2f} (1: {np.mean(rounds == 1):.2f}, 2: {np.mean(rounds == 2):.2f}, >2: {np.mean(rounds > 2):.2f});"
      f" tiles that polled an INVALID status word: {np.mean(polls > 0):.3f}, polls per tile (mean over lanes and tiles) {polls.mean():.2f}")
by = {}
for lo, hi in ((0, 80), (80, 160), (160, 230), (230, 296)):
    m = (cur[:, 10] >= lo) & (cur[:, 10] < hi)
    if m.any():
        print(f"    CTAs {lo:3d}-{hi - 1:3d}: resolve {med((cur[:, 6] - nx[:, 5])[m]):5.2f} us, window wait {med((cur[:, 12] - nx[:, 5])[m]):5.2f}, walk {med((cur[:, 6] - cur[:, 12])[m]):5.2f},"
              f" rounds {rounds[m].mean():.2f}, polled {np.mean(polls[m] > 0):.3f}")

# ---- skew between CTAs: when does a tile publish its aggregate relative to the others of its generation, and how long
# before (negative: AFTER) the start of its successor's resolve was each of the 160 predecessors' aggregate published?
pub = st[:, 5]
g0 = 50
gen = pub[g0 * grid:(g0 + 1) * grid] - pub[g0 * grid:(g0 + 1) * grid].min()
q = np.percentile(gen, [0, 10, 50, 90, 100]) / 1e3
print(f"  generation {g0}: publish time of the {grid} tiles relative to the first: p0 {q[0]:.2f} p10 {q[1]:.2f} p50 {q[2]:.2f} p90 {q[3]:.2f} p100 {q[4]:.2f} us;"
      f" median first half (CTAs 0-147) {np.median(gen[:148]) / 1e3:.2f}, second half {np.median(gen[148:]) / 1e3:.2f}")
ts = t[ok]
res_start = nx[:, 5]           # resolve of tile t starts right after its successor's publish (same warp, same iteration)
for d in (1, 2, 4, 8, 16, 32, 64, 128, 160):
    slack = (res_start - pub[ts - d]) / 1e3
    print(f"    predecessor t-{d:3d}: published {np.median(slack):6.2f} us before the resolve starts (p10 {np.percentile(slack, 10):6.2f}, p1 {np.percentile(slack, 1):6.2f}); not yet: {np.mean(slack < 0):.3f}")

# ---- who paces the kernel?  Per CTA: share of tiles that had to poll, and the phases of the CTAs that (almost) never wait
cta_ok = cur[:, 10]
frac = np.array([np.mean(polls[cta_ok == b] > 0) if np.any(cta_ok == b) else np.nan for b in range(grid)])
order = np.argsort(frac)
print(f"  per-CTA share of tiles that polled: min {np.nanmin(frac):.2f} p10 {np.nanpercentile(frac, 10):.2f} p50 {np.nanpercentile(frac, 50):.2f} p90 {np.nanpercentile(frac, 90):.2f} max {np.nanmax(frac):.2f}")
def phases(mask, label):
    print(f"    {label}: period {med((nx[:, 1] - cur[:, 1])[mask]):.2f}  data wait {med((cur[:, 2] - cur[:, 1])[mask]):.2f}  evaluate (worker) {med((cur[:, 11] - cur[:, 1])[mask]):.2f}"
          f"  totals+publish {med((cur[:, 5] - cur[:, 4])[mask]):.2f}  resolve {med((cur[:, 6] - nx[:, 5])[mask]):.2f}  output {med((cur[:, 8] - cur[:, 7])[mask]):.2f}"
          f"  start->b1 exit (thread 0) {med((cur[:, 4] - cur[:, 1])[mask]):.2f}  b1 exit -> next start {med((nx[:, 1] - cur[:, 4])[mask]):.2f}")
pacers = order[:15]
waiters = order[-15:]
phases(np.isin(cta_ok, pacers), f"15 CTAs that poll least {sorted(pacers.tolist())}")
phases(np.isin(cta_ok, waiters), f"15 CTAs that poll most  {sorted(waiters.tolist())}")
phases(polls == 0, "tiles without a poll")
phases(polls > 0, "tiles with a poll   ")

# ---- per CTA: its own work per tile (no waiting for others) vs the time it spends waiting in the resolve
own = (cur[:, 2] - cur[:, 1]) + (cur[:, 11] - cur[:, 1]) + (cur[:, 5] - cur[:, 4]) + (cur[:, 8] - cur[:, 7])   # data wait + evaluate + totals/publish + output
walkt = cur[:, 6] - cur[:, 12]
percta = []
for b in range(grid):
    m = cta_ok == b
    if m.any(): percta.append((b, float(np.mean(own[m])) / 1e3, float(np.mean(walkt[m])) / 1e3, float(np.mean(polls[m] > 0)), float(np.mean((cur[:, 2] - cur[:, 1])[m])) / 1e3, float(np.mean((cur[:, 8] - cur[:, 7])[m])) / 1e3))
percta.sort(key=lambda r: -r[1])
print("  CTAs with the most OWN work per tile (data wait + evaluate + totals + output, us) / walk / polled share / data wait / output:")
for r in percta[:12]: print(f"    CTA {r[0]:3d}: own {r[1]:.2f}  walk {r[2]:.2f}  polled {r[3]:.2f}  data {r[4]:.2f}  output {r[5]:.2f}")
print("  ... and the least:")
for r in percta[-6:]: print(f"    CTA {r[0]:3d}: own {r[1]:.2f}  walk {r[2]:.2f}  polled {r[3]:.2f}  data {r[4]:.2f}  output {r[5]:.2f}")
o = np.array([r[1] for r in percta]); w = np.array([r[2] for r in percta])
print(f"  own work per tile over the CTAs: min {o.min():.2f} p50 {np.median(o):.2f} max {o.max():.2f} us; correlation(own work, walk time) = {np.corrcoef(o, w)[0, 1]:.2f}")
