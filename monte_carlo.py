"""BASELINE.json configs[4]: a ~200-op fused Monte-Carlo mega-trace (PCG RNG, transcendentals, masked
select) over 2^26 lanes, built through the vkjit Python front-end.

`build(M, n, rounds)` only uses the front-end surface (`arange`, operators on `Var`, `sqrt/log/sin/cos/
exp`, `maximum`, `select`), so the same function drives the product (`vkjit_b200.vkjit`) and, in the
parity test, an adapter over the CPU oracle.  No input arrays: 4 B/lane written, ~215 lane-ops/lane —
bound by SM issue rate (FP32/INT32 + MUFU), not HBM.
"""
import time

SEED = 0xB2000001 + 5 * 16
U32, F32 = 3, 5


def pcg_out(s):
    w = ((s >> ((s >> 28) + 4)) ^ s) * 277803737
    return (w >> 22) ^ w


def build(M, n, rounds=5):
    lane = M.arange(U32, n)
    s = pcg_out((lane ^ SEED) * 747796405 + 2891336453)        # per-lane seed = pcg_hash(lane ^ SEED)
    acc = None
    for _ in range(rounds):
        s = s * 747796405 + 2891336453
        h1 = pcg_out(s)
        s = s * 747796405 + 2891336453
        h2 = pcg_out(s)
        u1 = ((h1 >> 8) + 1).cast(F32) * (2.0 ** -24)           # (0, 1]
        u2 = (h2 >> 8).cast(F32) * (2.0 ** -24)                 # [0, 1)
        rad = M.sqrt(M.log(u1) * -2.0)                          # Box-Muller
        th = u2 * 6.2831854820251465
        z0, z1 = rad * M.cos(th), rad * M.sin(th)
        price = M.exp(z0 * 0.2 + 0.01) * 100.0                  # one log-normal step
        pay = M.maximum(price - 100.0, 0.0)                     # call payoff
        term = M.select(z1 > 0.0, pay, pay * 0.5)               # masked select on the second normal
        acc = term if acc is None else acc + term
    return acc * (1.0 / rounds)


def bench(vk, stream, flush_l2, log2n=26, rounds=5):
    import torch

    from vkjit_b200 import vkjit
    n = 1 << log2n
    vk.cache_clear()
    vk.stats_reset()
    t0 = time.perf_counter()
    y = build(vkjit, n, rounds)
    t_trace = time.perf_counter() - t0
    text = vkjit.ir()
    nodes = text.count("Var {") - text.count("op: Free")
    t0 = time.perf_counter()
    vkjit.eval([y])
    vk.sync()
    t_cold = time.perf_counter() - t0
    st = vk.stats()
    compile_ms = st["last_compile_ns"] / 1e6
    times, evals = [], []
    for i in range(6):
        y = build(vkjit, n, rounds)
        flush_l2()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        vkjit.eval([y])
        b.record(stream)
        evals.append(vk.stats()["last_eval_ns"])
        vk.sync()
        if i >= 1:
            times.append(a.elapsed_time(b))
    ms = sum(times) / len(times)
    st2 = vk.stats()
    del y
    # cached-trace launch cost in steady state: the same 364-node program at n = 1024 hits the same cached kernel (n is a
    # kernel parameter), so 300 build + eval rounds run back to back without a 1 ms kernel or an L2 flush in between
    steady = []
    for i in range(300):
        y = build(vkjit, 1024, rounds)
        vkjit.eval([y])
        steady.append(vk.stats()["last_eval_ns"])
    vk.sync()
    st3 = vk.stats()
    del y
    steady = sorted(steady[20:])
    return {"lanes": n, "rounds": rounds, "ir_nodes_live_after_build": nodes, "trace_build_ms_python": t_trace * 1e3,
            "compile_ms_cold": compile_ms, "first_eval_wall_ms": t_cold * 1e3, "kernel_ms": ms, "Glanes_per_s": n / (ms * 1e-3) / 1e9,
            "cache_hit_eval_us": sorted(evals[1:])[len(evals[1:]) // 2] / 1e3, "cache_hits": st2["cache_hits"], "cache_misses": st2["cache_misses"],
            "cache_hit_eval_us_steady": {"median": steady[len(steady) // 2] / 1e3, "p90": steady[int(len(steady) * 0.9)] / 1e3,
                                         "evals": 300, "n": 1024, "cache_misses": st3["cache_misses"] - st2["cache_misses"]},
            "GBps_written": 4 * n / (ms * 1e-3) / 1e9,
            # lane-ops/s against 148 SMs x 128 FP32/INT32 lanes x clock.  One IR node is one lane-op here,
            # although sin/cos/log/sqrt/div expand to tens of instructions each, so the fraction understates ALU use.
            "T_ir_ops_per_s": nodes * n / (ms * 1e-3) / 1e12, "alu_peak_T_lane_ops_per_s": 148 * 128 * 1.965e9 / 1e12,
            "alu_frac_by_ir_ops": nodes * n / (ms * 1e-3) / (148 * 128 * 1.965e9)}
