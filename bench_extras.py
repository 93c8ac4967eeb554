"""Secondary measurements reported under "extras" in bench.py's JSON line (N=1 only).

The other BASELINE.json configs, each against the roofline that bounds it:
  E28  fused elementwise z = x*y + c over 2^28 f32      12 B/lane   HBM
  E20  arange*y + c over 2^20 f32 with readback          latency (config 0 shape)
  LAUNCH  cached-trace launch overhead, n = 1024          host us per eval
  C28  prefix sum / compress over 2^28 u32               8 / ~10 B/lane  HBM
  H26  gather + scatter-add, 2^26 indices -> 2^16 bins   atomic throughput
  M26  ~200-op Monte-Carlo trace over 2^26 lanes         SM issue rate; compile ms
"""
import time

import numpy as np
import torch

from bench import hash_trace, uniform_trace
from vkjit_b200.ir import Bop, Red, VarType as T


def _timed(stream, flush_l2, sync, fn, reps=5, warm=2):
    ts, last = [], None
    for i in range(warm + reps):
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        last = fn()
        b.record(stream)
        sync()
        if i >= warm:
            ts.append(a.elapsed_time(b))
    # median, not mean: one outlier among 3-5 repetitions (seen once in round 2: 0.93 ms reported for a 0.33 ms kernel)
    # must not become the figure; the best time is reported next to it
    ts.sort()
    return ts[len(ts) // 2], ts[0], last


def run(ir, vk, stream, flush_l2, peak):
    out = {}
    sync = vk.sync
    c = ir.const_u32

    # ---------------- E28: z = x*y + c
    n = 1 << 28
    lanes = ir.arange(T.U32, n)
    x = uniform_trace(ir, lanes, 0xB2000001 + 1 * 16 + 0)
    y = uniform_trace(ir, lanes, 0xB2000001 + 1 * 16 + 1)
    ir.eval([x, y])
    half = ir.const_f32(0.5)

    def e28():
        z = ir.add(ir.mul(x, y), half)
        ir.eval([z])
        ir.dec_ref_count(z)

    ms, best, _ = _timed(stream, flush_l2, sync, e28)
    gbs = 12 * n / (ms * 1e-3) / 1e9
    out["E28_fused_elementwise"] = {"ms": ms, "best_ms": best, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": 12, "n": n}

    # fused trace -> reduce: sum(x*y + c) in ONE generated kernel (z is never materialised: 8 B/lane)
    def e28r():
        z = ir.add(ir.mul(x, y), half)
        s = ir.reduce(Red.Sum, z)
        ir.dec_ref_count(z); ir.dec_ref_count(s)

    ms, best, _ = _timed(stream, flush_l2, sync, e28r, reps=3)
    out["E28_then_sum"] = {"ms": ms, "GBps_algorithmic_8B": 8 * n / (ms * 1e-3) / 1e9}

    # ---------------- C28: prefix sum and compress over 2^28 u32
    vals = hash_trace(ir, lanes, 0xB2000001 + 4 * 16 + 0)
    ir.eval([vals])

    def scan():
        r = ir.prefix_sum(vals, True)
        ir.dec_ref_count(r)

    ms, best, _ = _timed(stream, flush_l2, sync, scan)
    gbs = 8 * n / (ms * 1e-3) / 1e9
    out["C28_prefix_sum"] = {"ms": ms, "best_ms": best, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": 8}
    ir.dec_ref_count(y)
    mask = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 0xB2000001 + 4 * 16 + 1), c(1)), c(0))
    ir.eval([mask])
    cnt = [0]

    def comp():
        r, k = ir.compress_values(vals, mask)
        cnt[0] = k
        ir.dec_ref_count(r)

    ms, best, _ = _timed(stream, flush_l2, sync, comp, reps=3)
    bpl = 8 + 4 * cnt[0] / n
    gbs = bpl * n / (ms * 1e-3) / 1e9
    out["C28_compress"] = {"ms_incl_count_readback": ms, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": bpl, "selected": cnt[0]}
    # fused-mask variant: the mask is a trace, computed inside the compaction kernel
    # (scan_fused.cuh): 4 B read (values) + 4p B written per lane; no mask array, no kernel that writes one
    def comp_fused():
        mk = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 0xB2000001 + 4 * 16 + 1), c(1)), c(0))
        r, k = ir.compress_values(vals, mk)
        cnt[0] = k
        ir.dec_ref_count(r); ir.dec_ref_count(mk)

    ms, best, _ = _timed(stream, flush_l2, sync, comp_fused, reps=3)
    bpl = 4 + 4 * cnt[0] / n
    gbs = bpl * n / (ms * 1e-3) / 1e9

    def mask_kernel():  # what the unfused route runs first: the elementwise kernel that writes the 4-byte mask
        mk = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 0xB2000001 + 4 * 16 + 1), c(1)), c(0))
        ir.eval([mk])
        ir.dec_ref_count(mk)

    mk_ms, _, _ = _timed(stream, flush_l2, sync, mask_kernel, reps=3)
    out["C28_compress_fused_mask"] = {"ms_incl_count_readback": ms, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": bpl,
                                      "selected": cnt[0], "unfused_route_ms": {"mask_kernel": mk_ms, "compress": out["C28_compress"]["ms_incl_count_readback"],
                                                                             "total": mk_ms + out["C28_compress"]["ms_incl_count_readback"]}}

    # threshold filter: compress_values(v, v > t) — the mask depends on the values themselves, which are streamed once
    def comp_thresh():
        mk = ir.gt(vals, c(0x80000000))
        r, k = ir.compress_values(vals, mk)
        cnt[0] = k
        ir.dec_ref_count(r); ir.dec_ref_count(mk)

    ms, best, _ = _timed(stream, flush_l2, sync, comp_thresh, reps=3)
    bpl = 4 + 4 * cnt[0] / n
    gbs = bpl * n / (ms * 1e-3) / 1e9
    out["C28_filter_gt_threshold"] = {"ms_incl_count_readback": ms, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": bpl, "selected": cnt[0]}

    # prefix sum of an unevaluated trace: nothing is read, 4 B/lane written
    def scan_fused():
        h = hash_trace(ir, lanes, 0xB2000001 + 4 * 16 + 0)
        r = ir.prefix_sum(h, True)
        ir.dec_ref_count(r); ir.dec_ref_count(h)

    ms, best, _ = _timed(stream, flush_l2, sync, scan_fused, reps=3)
    gbs = 4 * n / (ms * 1e-3) / 1e9
    out["C28_prefix_sum_of_trace"] = {"ms": ms, "GBps": gbs, "hbm_frac": gbs / peak, "bytes_per_lane": 4}
    ir.dec_ref_count(mask); ir.dec_ref_count(vals); ir.dec_ref_count(x)

    # ---------------- H26: gather + scatter-add histogram
    m = 1 << 26
    lanes26 = ir.arange(T.U32, m)
    idx = ir.bop(Bop.And, hash_trace(ir, lanes26, 0xB2000001 + 3 * 16 + 0), c(0xFFFF))
    table = hash_trace(ir, ir.arange(T.U32, 1 << 16), 0xB2000001 + 3 * 16 + 1)
    ir.eval([idx])
    ir.eval([table])
    bins = ir.array_u32(np.zeros(1 << 16, np.uint32))

    def hist():
        w = ir.gather(table, idx)
        s = ir.scatter_add(w, bins, idx)
        ir.eval([s])
        ir.dec_ref_count(s)

    ms, best, _ = _timed(stream, flush_l2, sync, hist)
    out["H26_gather_scatter_add"] = {"ms": ms, "best_ms": best, "Gelem_per_s": m / (ms * 1e-3) / 1e9,
                                     "GBps_algorithmic_4B": 4 * m / (ms * 1e-3) / 1e9, "bins": 1 << 16,
                                     "bound": "atomics (49152 of 65536 bins privatised in shared memory, rest L2 RED)"}
    one = ir.const_u32(1)

    def count():
        s = ir.scatter_add(one, bins, idx)
        ir.eval([s])
        ir.dec_ref_count(s)

    ms, best, _ = _timed(stream, flush_l2, sync, count)
    out["H26_count_histogram"] = {"ms": ms, "best_ms": best, "Gelem_per_s": m / (ms * 1e-3) / 1e9, "bins": 1 << 16,
                                  "bound": "shared-memory + L2 atomics"}
    ir.dec_ref_count(idx)

    # skewed indices: min of two uniform draws — low bins are hit about twice as often
    hh = hash_trace(ir, lanes26, 0xB2000001 + 3 * 16 + 0)
    idx_s = ir.bop(Bop.Min, ir.bop(Bop.And, hh, c(0xFFFF)), ir.shr(hh, c(16)))
    ir.eval([idx_s])

    def hist_skew():
        w = ir.gather(table, idx_s)
        s = ir.scatter_add(w, bins, idx_s)
        ir.eval([s])
        ir.dec_ref_count(s)

    ms, best, _ = _timed(stream, flush_l2, sync, hist_skew)
    out["H26_skewed_gather_scatter_add"] = {"ms": ms, "best_ms": best, "Gelem_per_s": m / (ms * 1e-3) / 1e9, "bins": 1 << 16,
                                            "indices": "min(h & 0xFFFF, h >> 16)"}
    ir.dec_ref_count(idx_s); ir.dec_ref_count(hh)

    # hot bins: 2^26 indices into 16 bins — every warp collides.  Shared-memory privatisation absorbs that; the
    # warp-aggregated variant (match.any + redux.sync, VKJIT_AGG=1) measured 6.7x SLOWER here (profiles/r02_h26_hot_bins.md)
    idx_h = ir.bop(Bop.And, hash_trace(ir, lanes26, 0xB2000001 + 3 * 16 + 2), c(15))
    ir.eval([idx_h])
    bins16 = ir.array_u32(np.zeros(16, np.uint32))

    def hist_hot():
        s = ir.scatter_add(one, bins16, idx_h)
        ir.eval([s])
        ir.dec_ref_count(s)

    ms, best, _ = _timed(stream, flush_l2, sync, hist_hot)
    import os as _os
    out["H26_hot_bins_count"] = {"ms": ms, "best_ms": best, "Gelem_per_s": m / (ms * 1e-3) / 1e9, "bins": 16,
                                 "warp_aggregation": "probed per warp (VKJIT_AGG=1)" if _os.environ.get("VKJIT_AGG") == "1" else "off (default: plain atomics into privatised bins)",
                                 "sum_of_bins_ok": int(ir.as_slice(bins16, T.U32).astype(np.uint64).sum()) % m == 0}
    ir.dec_ref_count(idx_h); ir.dec_ref_count(bins16)

    # ---------------- E20 with readback + cached launch overhead
    n20 = 1 << 20
    y20 = uniform_trace(ir, ir.arange(T.U32, n20), 7)
    ir.eval([y20])
    host = np.empty(n20, np.float32)
    t = []
    for i in range(12):
        t0 = time.perf_counter()
        z = ir.add(ir.mul(ir.arange(T.F32, n20), y20), half)
        ir.eval([z])
        ir.read_into(z, T.F32, host.ctypes.data, host.nbytes)
        t.append(time.perf_counter() - t0)
        ir.dec_ref_count(z)
    out["E20_eval_plus_readback_us"] = {"median": 1e6 * sorted(t[2:])[len(t[2:]) // 2], "n": n20, "readback_bytes": 4 * n20,
                                        "host_buffer": "caller's pageable buffer (vkjit_read into a numpy array): through the pinned staging ring"}
    t = []
    for i in range(12):   # the DEFAULT front-end read: as_slice returns an array that lives in pinned memory (one DMA, no staging copy)
        t0 = time.perf_counter()
        z = ir.add(ir.mul(ir.arange(T.F32, n20), y20), half)
        ir.eval([z])
        res = ir.as_slice(z, T.F32)
        t.append(time.perf_counter() - t0)
        ir.dec_ref_count(z)
    out["E20_eval_plus_readback_default_us"] = {"median": 1e6 * sorted(t[2:])[len(t[2:]) // 2], "n": n20, "readback_bytes": 4 * n20,
                                                "host_buffer": "default: Ir.as_slice / Var.numpy() result array from the pinned pool",
                                                "checksum": float(res[:16].sum())}
    del res
    import ctypes as C
    hp = C.c_void_p()
    vk.product_api().call("host_alloc", 4 * n20, C.byref(hp))   # pinned staging (vkjit_host_alloc)
    t = []
    for i in range(12):
        t0 = time.perf_counter()
        z = ir.add(ir.mul(ir.arange(T.F32, n20), y20), half)
        ir.eval([z])
        ir.read_into(z, T.F32, hp.value, 4 * n20)
        t.append(time.perf_counter() - t0)
        ir.dec_ref_count(z)
    vk.product_api().call("host_free", hp)
    out["E20_eval_plus_readback_pinned_us"] = {"median": 1e6 * sorted(t[2:])[len(t[2:]) // 2], "n": n20, "readback_bytes": 4 * n20,
                                               "host_buffer": "pinned (vkjit_host_alloc)"}

    a1k = ir.arange(T.F32, 1024)
    sync()
    reps = 3000
    evals = []
    vk.stats_reset()
    t0 = time.perf_counter()
    for i in range(reps):
        z = ir.add(ir.mul(a1k, half), half)
        ir.eval([z])
        evals.append(vk.stats()["last_eval_ns"]) if i % 50 == 0 else None
        ir.dec_ref_count(z)
    host_us = (time.perf_counter() - t0) / reps * 1e6
    sync()
    st = vk.stats()
    out["LAUNCH_cached_trace"] = {"python_loop_us_per_iter": host_us, "vkjit_eval_us_median": sorted(evals)[len(evals) // 2] / 1e3,
                                  "cache_hits": st["cache_hits"], "cache_misses": st["cache_misses"], "n": 1024}
    # the same loop through the `vkjit` Python module (native Var type, csrc/pyfront.cpp): what a front-end user pays
    from vkjit_b200 import vkjit as vj
    b1k = vj.arange(T.F32, 1024)
    vj.eval([b1k * 0.5 + 0.5])
    sync()
    t0 = time.perf_counter()
    for i in range(reps):
        z = b1k * 0.5 + 0.5
        vj.eval([z])
    out["LAUNCH_cached_trace"]["vkjit_module_loop_us_per_iter"] = (time.perf_counter() - t0) / reps * 1e6
    sync()
    del z, b1k
    try:
        import monte_carlo
        out["M26_monte_carlo"] = monte_carlo.bench(vk, stream, flush_l2)
    except ImportError:
        pass
    return out
