#!/usr/bin/env python
"""Per-launch timeline of the sharded reduction chain (VERDICT r01 task 1b): where do the microseconds of one
reduction go on every rank — streaming, waiting for the previous kernel, ticket, fold, peer exchange — and how long
is the launch-to-launch period against the streaming time?

    VKJIT_REDUCE_TRACE=1 torchrun --nproc-per-node N profiles/reduce_timeline.py --steps 20 [--out gpurun_out/timeline]

Uses the %globaltimer stamps of prims.cu: reduce_kernel (include/vkjit_b200.h: vkjit_debug_reduce_trace).  Stamps are
only compared within one GPU.  Writes <out>_rank<r>.json and, on rank 0, a summary table to stdout.
"""
import argparse
import ctypes as C
import json
import os
import statistics
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("VKJIT_REDUCE_TRACE", "1")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--log2n", type=int, default=28)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline"))
    ap.add_argument("--no-align", action="store_true", help="skip the in-stream alignment before the chain (round-1 behaviour)")
    ap.add_argument("--arrays", type=int, default=4, help="number of input arrays the chain rotates over")
    args = ap.parse_args()
    import torch
    import torch.distributed as td

    import vkjit_b200 as vk
    from bench import SEED_R28, uniform_trace
    from vkjit_b200.ir import Ir, Red, VarType as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    vk.init(local)
    api = vk.product_api()
    if world > 1:
        from vkjit_b200 import dist
        dist.init_from_torch(dev, p2p=True)
    ir = Ir()
    n = 1 << args.log2n
    lanes = ir.arange_sharded(T.U32, n)
    xs = [uniform_trace(ir, lanes, SEED_R28 + i) for i in range(args.arrays)]
    for v in xs:
        ir.eval([v])
    tiny = ir.cast(ir.arange_sharded(T.U32, 4096 * world), T.F32) if world > 1 else None
    if tiny is not None:
        ir.eval([tiny])
    vk.sync()

    def barrier():
        vk.sync(); torch.cuda.synchronize()
        if world > 1:
            td.barrier(); torch.cuda.synchronize()

    def chain(k):
        out = []
        for i in range(k):
            out.append(ir.reduce(Red.Sum, xs[(2 * i) % len(xs)]))
            out.append(ir.reduce(Red.Max, xs[(2 * i + 1) % len(xs)]))
        return out

    for v in chain(5):
        ir.dec_ref_count(v)
    barrier()
    buf = (C.c_uint64 * (8 * 4096))()
    cnt = C.c_size_t()
    api.call("debug_reduce_trace", buf, len(buf), C.byref(cnt))   # clear
    barrier()
    if tiny is not None and not args.no_align:
        for _ in range(2):
            ir.dec_ref_count(ir.reduce(Red.Sum, tiny))
    t0 = time.perf_counter()
    res = chain(args.steps)
    t_issue = time.perf_counter() - t0
    vk.sync()
    t_all = time.perf_counter() - t0
    api.call("debug_reduce_trace", buf, len(buf), C.byref(cnt))
    rows = [[int(buf[i * 8 + j]) for j in range(8)] for i in range(cnt.value)]
    rows = rows[-2 * args.steps:]
    for v in res:
        ir.dec_ref_count(v)

    def us(a, b):
        return (b - a) / 1e3

    rec = []
    for i, r in enumerate(rows):
        d = {"stream_us": us(r[0], r[1]), "wait_prev_us": us(r[1], r[2]), "ticket_us": us(r[2], r[3]), "fold_us": us(r[3], r[4]),
             "exchange_us": us(r[4], r[5]), "kernel_us": us(r[0], r[5])}
        if i + 1 < len(rows):
            d["period_us"] = us(r[0], rows[i + 1][0])
            d["next_start_minus_done_us"] = us(r[5], rows[i + 1][0])
        rec.append(d)
    summary = {"rank": rank, "world": world, "n_local": ir.size(xs[0]), "launches": len(rows), "host_issue_us_per_reduction": t_issue / (2 * args.steps) * 1e6,
               "wall_us_per_reduction": t_all / (2 * args.steps) * 1e6,
               "chain_us": us(rows[0][0], rows[-1][5]) if rows else None}
    for key in ("stream_us", "wait_prev_us", "ticket_us", "fold_us", "exchange_us", "kernel_us", "period_us", "next_start_minus_done_us"):
        vals = [d[key] for d in rec[1:] if key in d]    # launch 0 carries the start-up skew
        if vals:
            summary[key] = {"median": statistics.median(vals), "mean": sum(vals) / len(vals), "max": max(vals), "min": min(vals)}
    summary["first_launch"] = rec[0] if rec else None
    # per input array: is the streaming time a property of where the array lives?
    per = {}
    for i, d in enumerate(rec):
        per.setdefault(i % len(xs), []).append(d["stream_us"])
    ptrs = []
    for v in xs:
        p = C.c_uint64()
        api.call("var_device_ptr", ir._h, v, C.byref(p))
        ptrs.append(p.value)
    summary["per_array"] = [{"array": a, "ptr": hex(ptrs[a]), "ptr_mod_128MiB_MiB": (ptrs[a] % (128 << 20)) / (1 << 20),
                             "stream_us_median": statistics.median(v[1:] or v), "stream_us_min": min(v)} for a, v in sorted(per.items())]
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(f"{args.out}_rank{rank}.json", "w") as f:
        json.dump({"summary": summary, "launches": rec}, f)
    gathered = [None] * world
    if world > 1:
        td.all_gather_object(gathered, summary)
    else:
        gathered = [summary]
    if rank == 0:
        print(json.dumps({"world": world, "steps": args.steps, "aligned": not args.no_align, "ranks": gathered}))
    barrier()
    ir.close()
    if world > 1:
        api.call("dist_shutdown")
        td.destroy_process_group()


if __name__ == "__main__":
    main()
