#!/usr/bin/env python
"""bench.py — the hot path measured on B200 (contract: one JSON line on rank 0).

Workload (BASELINE.json configs[1]): horizontal sum + max over 2^28 f32, contiguously sharded
over the N GPUs of one box, per-GPU partials combined by an all-reduce over NVLink.
One step = sum(x) and max(x) over the whole 2^28-lane array = 2 GiB of algorithmic reads.

  python bench.py [--gpus N --steps K --warmup W]       our CUDA path (through the C ABI)
  python bench.py --impl reference ...                   the CPU restatement of the reference
                                                         (oracle port) on the box's host cores

Timing: the K steps run back to back between ONE pair of CUDA events on the backend stream (barrier +
synchronize on both sides), max over ranks.  The steps rotate over four independent input arrays
(4 GiB/N per GPU, far beyond the 126 MB L2; profiles/rotation_ab.py shows that 1, 2, 4 or 8 arrays give
the same time, i.e. nothing is served from L2), so no eviction kernel sits inside the timed region.
The per-reduction figure with an L2 eviction and an event pair around every single reduction (the
method of the first bench lines of this round) is kept as `isolated`.  Clocks are sampled through NVML
during the timed region.  `e2e` goes through the same API with pinned HOST buffers: H2D of the step's
input, both reductions, D2H of the two results, wall clock.
"""
import argparse
import ctypes as C
import json
import math
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG2N = 28
N_TOTAL = 1 << LOG2N
SEED_R28 = 0xB2000001 + 2 * 16  # 0xB200_0001 + config*16 + array#
METRIC = "fused elementwise/reduce HBM GB/s (sum+max over 2^28 f32, sharded)"
UNIT = "GB/s"
# the SAME dict in both arms (the driver compares them); everything arm-specific lives under "setup"
CONFIG = {"workload": "R28: sum+max over 2^28 f32 (BASELINE.json configs[1])", "n": N_TOTAL,
          "l2": "inputs larger than L2 (every reduction reads 1 GiB / N per GPU, a different array than the one before)"}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md, MEASURED_PEAKS.json absent)"


# ------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML (what nvidia-smi reads) every ~1 ms in a background
    thread DURING the timed region (the region is only tens of ms long, too short for `nvidia-smi -lms`).
    NVML is initialised up front (prepare) so that the thread samples from the first millisecond; one sample is
    also taken synchronously at start and at stop."""

    def __init__(self, device):
        self.device, self.samples, self.reasons, self.max_mhz = device, [], set(), None
        self._stop = threading.Event()
        self._thr, self._nv, self._h, self.error = None, None, None, None
        self.period_s = float(os.environ.get("BENCH_CLOCK_PERIOD_MS", "1")) * 1e-3

    def prepare(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            self._nv, self._h = nv, nv.nvmlDeviceGetHandleByIndex(self.device)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._h, nv.NVML_CLOCK_SM))
            self._bits = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                          "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        except Exception as ex:  # pragma: no cover
            self.error = repr(ex)

    def _sample(self):
        nv = self._nv
        self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self._h)
        for k, b in self._bits.items():
            if r & b:
                self.reasons.add(k)

    def _run(self):
        try:
            while not self._stop.is_set():
                self._sample()
                time.sleep(self.period_s)
        except Exception as ex:  # pragma: no cover
            self.error = repr(ex)

    def start(self):
        if self._nv is None:
            self.prepare()
        if self._nv is None:
            return
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=5)
        try:
            if self._nv is not None:
                self._sample()   # one more sample right at the end of the region (GPU still warm)
        except Exception as ex:  # pragma: no cover
            self.error = repr(ex)
        out = {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.samples), "how": "NVML, 1 ms period, during the timed region"}
        if self.error:
            out["error"] = self.error
        return out


# ------------------------------------------------------------------------------------------
# synthetic input: the stateless generator of SURVEY.md §8d, built as a vkjit trace
# ------------------------------------------------------------------------------------------
def hash_trace(ir, lane_u32, seed):
    from vkjit_b200.ir import Bop
    c = ir.const_u32
    x = ir.bop(Bop.Xor, lane_u32, c(seed))
    s = ir.add(ir.mul(x, c(747796405)), c(2891336453))
    sh = ir.add(ir.bop(Bop.Shr, s, c(28)), c(4))
    w = ir.mul(ir.bop(Bop.Xor, ir.bop(Bop.Shr, s, sh), s), c(277803737))
    return ir.bop(Bop.Xor, ir.bop(Bop.Shr, w, c(22)), w)


def uniform_trace(ir, lane_u32, seed):
    from vkjit_b200.ir import Bop, VarType as T
    h = hash_trace(ir, lane_u32, seed)
    return ir.mul(ir.cast(ir.bop(Bop.Shr, h, ir.const_u32(8)), T.F32), ir.const_f32(2.0 ** -24))


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (cpu_baseline leg and --impl reference)
# ------------------------------------------------------------------------------------------
def cpu_reduce_arm(steps, warmup, log2n=LOG2N):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import oracle_api
    from vkjit_b200.ir import Ir, Red, VarType as T
    api = oracle_api()
    try:
        cores = len(os.sched_getaffinity(0))   # the cores this process may use (cgroup / taskset), not the box's total
    except AttributeError:
        cores = os.cpu_count() or 1
    api.call("set_threads", cores)            # worker t is pinned to the t-th allowed core (oracle.cpp: parallel_for)
    n = 1 << log2n
    ir = Ir(_api=api)
    x = ir.array_empty(T.F32, n)
    p = C.c_void_p()
    api.call("var_host_ptr", ir._h, x, C.byref(p))
    api.call("fill_hash", p, n, 0, SEED_R28, 1)
    times, res = [], None
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        s, m = ir.reduce(Red.Sum, x), ir.reduce(Red.Max, x)
        dt = time.perf_counter() - t0
        res = (float(ir.as_slice(s, T.F32)[0]), float(ir.as_slice(m, T.F32)[0]))
        ir.dec_ref_count(s); ir.dec_ref_count(m)
        if it >= warmup:
            times.append(dt)
    ir.close()
    t = sum(times) / len(times)
    return {"value": 2 * n * 4 / t / 1e9, "seconds_per_step": t, "cores": cores, "n": n, "sum": res[0], "max": res[1]}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_reduce_arm(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": dict(CONFIG),
        "setup": {"note": "CPU restatement of the reference semantics (oracle port), NOT the reference's Vulkan backend on Mesa "
                          "lavapipe: no Rust/Vulkan toolchain exists in this image",
                  "threads": r["cores"], "pinned": True},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": "full workload: sum+max over 2^28 f32 per step, all host threads (pinned)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "result": {"sum": r["sum"], "max": r["max"]},
    }
    print(json.dumps(line), flush=True)


def cpu_extras_arm():
    """CPU legs of the secondary configs (BASELINE.json configs 0, 2-4 and the north_star's E28): the oracle port on all
    host cores (pinned), each on a BOUNDED sample of the workload (the sample is stated; rates are per lane, so they
    carry over).  Printed as one JSON line by `bench.py --impl reference --extras`; bench_extras attaches them as
    `cpu_baseline` next to each GPU figure."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import numpy as np
    from oracle_lib import OracleIr, oracle_api
    from vkjit_b200.ir import Bop, Red, VarType as T
    api = oracle_api()
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    api.call("set_threads", cores)
    out = {}

    def timed(fn, reps=2):
        fn()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return min(ts)

    def base(value, unit, sample):
        return {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample}

    ir = OracleIr()
    c = ir.const_u32
    half = ir.const_f32(0.5)
    # E20 (configs[0]): arange * y + c over 2^20 f32 with readback — the full workload
    n20 = 1 << 20
    y20 = uniform_trace(ir, ir.arange(T.U32, n20), 7); ir.eval([y20])

    def e20():
        z = ir.add(ir.mul(ir.arange(T.F32, n20), y20), half); ir.eval([z]); ir.as_slice(z, T.F32); ir.dec_ref_count(z)
    out["E20_eval_plus_readback_us"] = base(timed(e20, 5) * 1e6, "us", "full workload (2^20 lanes, eval + readback)")
    # E28: z = x*y + c, sample 2^26 of 2^28 lanes
    n = 1 << 26
    lanes = ir.arange(T.U32, n)
    x, y = uniform_trace(ir, lanes, 17), uniform_trace(ir, lanes, 18); ir.eval([x, y])

    def e28():
        z = ir.add(ir.mul(x, y), half); ir.eval([z]); ir.dec_ref_count(z)
    out["E28_fused_elementwise"] = base(12 * n / timed(e28) / 1e9, "GB/s", "2^26 of 2^28 lanes (12 B/lane)")
    ir.dec_ref_count(y)
    # C28: prefix sum and compress, sample 2^26 of 2^28
    vals = hash_trace(ir, lanes, 65); ir.eval([vals])
    out["C28_prefix_sum"] = base(8 * n / timed(lambda: ir.dec_ref_count(ir.prefix_sum(vals, True))) / 1e9, "GB/s", "2^26 of 2^28 lanes (8 B/lane)")
    mask = ir.neq(ir.bop(Bop.And, hash_trace(ir, lanes, 66), c(1)), c(0)); ir.eval([mask])

    def comp():
        r, _k = ir.compress_values(vals, mask); ir.dec_ref_count(r)
    out["C28_compress"] = base(10 * n / timed(comp) / 1e9, "GB/s", "2^26 of 2^28 lanes (10 B/lane at p = 0.5)")
    ir.dec_ref_count(mask); ir.dec_ref_count(vals); ir.dec_ref_count(x)
    # H26: gather + scatter-add, sample 2^24 of 2^26 indices into 2^16 bins
    m = 1 << 24
    idx = ir.bop(Bop.And, hash_trace(ir, ir.arange(T.U32, m), 49), c(0xFFFF)); ir.eval([idx])
    table = hash_trace(ir, ir.arange(T.U32, 1 << 16), 50); ir.eval([table])
    bins = ir.array_u32(np.zeros(1 << 16, np.uint32))

    def hist():
        s = ir.scatter_add(ir.gather(table, idx), bins, idx); ir.eval([s]); ir.dec_ref_count(s)
    out["H26_gather_scatter_add"] = base(m / timed(hist) / 1e9, "Gelem/s", "2^24 of 2^26 indices, 2^16 bins")
    ir.close()
    # M26: the Monte-Carlo mega-trace, sample 2^22 of 2^26 lanes
    import monte_carlo
    from ir_adapter import IrModule
    ir = OracleIr()
    nm = 1 << 22

    def mc():
        yv = monte_carlo.build(IrModule(ir), nm, 5); ir.eval([yv.id])
    out["M26_monte_carlo"] = base(nm / timed(mc) / 1e9, "Glanes/s", "2^22 of 2^26 lanes (364-node trace, 5 rounds)")
    ir.close()
    print(json.dumps({"impl": "reference", "extras_cpu_baseline": out}), flush=True)


def cpu_arm_subprocess(steps=5, warmup=2):
    """The CPU leg of OUR arm is the reference arm itself, run as a child process with the same code, thread count and
    pinning (`bench.py --impl reference`): the two figures cannot drift apart (round 1: 37.8 vs 57.3 GB/s on one box,
    because the in-process leg shared the host with torch / NCCL helper threads and timed two passes only)."""
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE")}
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", str(steps), "--warmup", str(warmup)],
                       capture_output=True, text=True, timeout=900, env=env, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    if r.returncode != 0 or not lines:
        raise RuntimeError("reference arm failed: " + r.stderr[-500:])
    return json.loads(lines[-1])


def mgpu_parity(world):
    """N>1: the sharded paths' parity test (tests/mgpu_worker.py: sharded elementwise, reduce sum/min/max, prefix sum,
    compress -> indices / values, fused and hand-written, P2P mailbox and NCCL, every rank against the oracle) run as
    `world` plain processes WITHOUT torch (vkjit_dist_init_env).  Rank 0 starts it after the timed legs are over and
    the job's own communicator is shut down, so that the multi-GPU run the driver does also executes the parity check
    (the driver's pytest box has one GPU).  The test lives in tests/; this only launches it and records the verdict."""
    t0 = time.perf_counter()
    port = int(os.environ.get("MASTER_PORT", "29500")) + 11
    procs = []
    for r in range(world):
        env = {k: v for k, v in os.environ.items() if not k.startswith("TORCHELASTIC") and k not in ("GROUP_RANK", "ROLE_RANK", "LOCAL_WORLD_SIZE")}
        env.update(RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), VKJIT_RDZV_PORT=str(port + 1))
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"), "--no-torch"], env=env, cwd=ROOT,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    ok, tail = True, ""
    for r, p in enumerate(procs):
        try:
            o, e = p.communicate(timeout=600)
        except subprocess.TimeoutExpired:
            p.kill(); o, e = p.communicate()
        if p.returncode != 0 or f"rank {r}/{world} ok" not in o:
            ok, tail = False, (o[-300:] + e[-700:])
    return {"ok": ok, "ranks": world, "seconds": time.perf_counter() - t0, "launcher": "plain processes, no torch (vkjit_dist_init_env)",
            "test": "tests/mgpu_worker.py: sharded elementwise / reduce / prefix sum / compress, P2P and NCCL, bit-exact vs the oracle",
            **({"error": tail} if not ok else {})}


# ------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------
def ours(args):
    import numpy as np
    import torch
    import torch.distributed as td

    import vkjit_b200 as vk
    from vkjit_b200.ir import Ir, Red, VarType as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        td.init_process_group("nccl", device_id=dev)
    vk.init(local_rank)
    api = vk.product_api()
    if world > 1:
        from vkjit_b200 import dist
        dist.init_from_torch(dev, p2p=True)   # NCCL unique id + cudaIpc mailbox handles through torch.distributed
        if args.collective == "nccl":
            dist.set_p2p(False)
    stream = torch.cuda.ExternalStream(vk.stream_ptr(), device=dev)
    ir = Ir()

    def barrier():
        # Drain this rank's stream BEFORE entering the NCCL barrier: the barrier's kernel runs on torch's stream, and
        # if it sits on the SMs next to a queue of back-to-back reductions it takes CTA slots away from them (a
        # reduction is exactly one full wave) until the slowest rank arrives — measured at N=2: 114 instead of 77 us
        # per reduction over the whole timed region.
        vk.sync()
        torch.cuda.synchronize()
        if world > 1:
            td.barrier()
            torch.cuda.synchronize()

    flush_buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)  # 256 MiB = 2x the 126 MB L2 (f32: summed in place, no upcast copy)

    def flush_l2():
        # Evict the inputs by streaming a 256 MiB READ through L2.  A write-flush would leave ~126 MB
        # of dirty lines whose write-back lands inside the next timed kernel (measured: +10 % on a
        # 1 GiB reduction, ncu shows the kernel itself at 154 us vs 173 us timed after a write-flush).
        with torch.cuda.stream(stream):
            flush_buf.sum()

    # ---- input: four independent arrays of 2^28 f32 uniform[0,1) (this rank's contiguous shard of each), generated
    # on the device.  Reduction j of the run reads array j mod 4, so an array is re-read only after 3 GiB/N of other
    # reads went through the 126 MB L2.
    ROT = 4
    lanes = ir.arange_sharded(T.U32, N_TOTAL)
    xs = [uniform_trace(ir, lanes, SEED_R28 + i) for i in range(ROT)]
    for v in xs:
        ir.eval([v])
    x = xs[0]
    n_local = ir.size(x)
    vk.sync()

    # N>1, isolated timing only: the L2-eviction kernels take a different time on every GPU; without re-alignment
    # that spread (~10 us) would be charged to the collective of the next timed reduction (each rank waits for the
    # slowest).  An UNTIMED tiny all-reduce after each eviction lines the streams up again.
    tiny = ir.cast(ir.arange_sharded(T.U32, 4096 * world), T.F32) if world > 1 else None
    if tiny is not None:
        ir.eval([tiny])

    def align():
        if tiny is not None:
            ir.dec_ref_count(ir.reduce(Red.Sum, tiny))

    pending = []

    def step(k):
        """One step: sum over one array, max over the next (2 GiB of algorithmic reads in total over all GPUs)."""
        s = ir.reduce(Red.Sum, xs[(2 * k) % ROT])
        m = ir.reduce(Red.Max, xs[(2 * k + 1) % ROT])
        pending.append((s, m))
        if len(pending) > 2:          # results are dropped two steps later: no host round trip in the loop
            for v in pending.pop(0):
                ir.dec_ref_count(v)

    # Warm-up: the same back-to-back chain as the timed region.
    for k in range(max(3, args.warmup)):
        step(k)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.prepare()   # NVML init outside the timed region
        sampler.start()     # the sampling thread starts BEFORE the barrier: nothing rank-specific sits between barrier and ev0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    vk.stats_reset()
    # N>1: the ranks leave the host-side barrier up to a few hundred us apart.  Without re-alignment that skew is charged to
    # the first exchange of the timed region (every rank's first tail waits for the last rank to arrive) — 0.4-0.6 ms
    # against a 1.5 ms region at N = 8, --steps 20 (round 1).  Two untimed tiny sharded reductions IN THE STREAM, right
    # before ev0, are a device-side barrier: every rank's ev0 fires within an NVLink round trip of the others'.
    align(); align()
    t_wall0 = time.perf_counter()
    ev0.record(stream)
    for k in range(args.steps):
        step(k)
    ev1.record(stream)
    t_issue = time.perf_counter() - t_wall0    # host time to put the K steps on the stream (no sync inside)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    st = vk.stats()
    clocks = sampler.stop() if rank == 0 else None
    step_ms = torch.tensor([ev0.elapsed_time(ev1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(step_ms, op=td.ReduceOp.MAX)
    ms_per_step = float(step_ms.item())
    total_bytes = 2 * N_TOTAL * 4
    value = total_bytes / (ms_per_step * 1e-3) / 1e9
    launches = st["trace_launches"] + st["prim_launches"] - (2 if world > 1 else 0)   # minus the two untimed alignment reductions
    for pr in pending:
        for v in pr:
            ir.dec_ref_count(v)
    pending.clear()
    # the result that is checked against the CPU side: sum and max of array 0 (untimed)
    s0, m0 = ir.reduce(Red.Sum, x), ir.reduce(Red.Max, x)
    gpu_sum = float(ir.as_slice(s0, T.F32)[0])
    gpu_max = float(ir.as_slice(m0, T.F32)[0])
    ir.dec_ref_count(s0); ir.dec_ref_count(m0)

    # ---- the same step timed reduction by reduction: L2 evicted (and, at N>1, ranks re-aligned) before each one,
    # one event pair per reduction.  Carries ~2.6 us of event overhead and the full launch latency per reduction.
    iso_steps = min(args.steps, 10)
    iso = []
    for i in range(2 + iso_steps):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        flush_l2(); align()
        evs[0].record(stream)
        s = ir.reduce(Red.Sum, xs[(2 * i) % ROT])
        evs[1].record(stream)
        flush_l2(); align()
        evs[2].record(stream)
        m = ir.reduce(Red.Max, xs[(2 * i + 1) % ROT])
        evs[3].record(stream)
        vk.sync()
        ir.dec_ref_count(s); ir.dec_ref_count(m)
        if i >= 2:
            iso.append((evs[0].elapsed_time(evs[1]), evs[2].elapsed_time(evs[3])))
    iso_ms = torch.tensor([sum(a + b for a, b in iso) / len(iso)], device=dev, dtype=torch.float64)
    if world > 1:
        td.all_reduce(iso_ms, op=td.ReduceOp.MAX)
    isolated = {"ms_per_step": float(iso_ms.item()), "value": total_bytes / (float(iso_ms.item()) * 1e-3) / 1e9, "unit": UNIT, "steps": iso_steps,
                "how": "L2 evicted by a 256 MiB read before every reduction, one CUDA-event pair per reduction (outside the headline)"}

    # ---- roofline of the dominant kernel (reduce_kernel<float,SUM>): kernel-only, no collective
    peak, peak_src = peaks()
    if world == 1:
        kern_ms = ms_per_step / 2          # two launches per step, timed live over the timed region
    else:
        xl = uniform_trace(ir, ir.arange(T.U32, n_local), SEED_R28)
        ir.eval([xl])
        ks = []
        for i in range(3 + args.steps):
            flush_l2()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); r = ir.reduce(Red.Sum, xl); b.record(stream)
            vk.sync(); ir.dec_ref_count(r)
            if i >= 3:
                ks.append(a.elapsed_time(b))
        kern_ms = sum(ks) / len(ks)
        ir.dec_ref_count(xl)
    achieved = n_local * 4 / (kern_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:  # DRAM traffic of this kernel at this size from the committed ncu capture (never measured under the bench)
        tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tr.get(f"reduce_kernel<float, SUM, 512> @ 2^{int(math.log2(n_local))} lanes") if n_local & (n_local - 1) == 0 else None
        if ent:
            traffic, traffic_src = ent["dram_read_bytes"] + ent["dram_write_bytes"], "profiles/r02_traffic.json (ncu --set full)"
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "reduce_kernel<float, SUM, 512>" + (" / <float, MAX, 512> (mean over the timed region's launches)" if world == 1 else " (kernel only, isolated)"), "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "launch_ms": kern_ms,
                "algorithmic_bytes_per_launch": n_local * 4}

    # ---- e2e: same metric through the public API, host <-> device copies inside the timed region.  Two legs:
    #   e2e                 the contract's definition: the step's input in PINNED host memory (vkjit_host_alloc)
    #   e2e_default_numpy   the default front-end path: an ordinary (pageable) NumPy array handed to array_*, which the
    #                       library moves through its pinned staging ring with a threaded host memcpy (csrc/staging.cpp)
    e2e_steps = max(2, min(args.steps, 5))
    hp = C.c_void_p()
    api.call("host_alloc", n_local * 4, C.byref(hp))
    ir.read_into(x, T.F32, hp.value, n_local * 4)       # the step input now lives in pinned host memory
    x_np = np.empty(n_local, np.float32)                # ... and in ordinary pageable memory
    C.memmove(x_np.ctypes.data, hp.value, n_local * 4)
    out2 = np.zeros(2, np.float32)

    def e2e_leg(upload):
        ts = []
        for i in range(1 + e2e_steps):
            barrier()
            t0 = time.perf_counter()
            xv = upload()                                                            # H2D of this step's input
            s2, m2 = ir.reduce(Red.Sum, xv), ir.reduce(Red.Max, xv)
            out2[0] = ir.as_slice(s2, T.F32)[0]; out2[1] = ir.as_slice(m2, T.F32)[0]  # D2H of the results
            dt = time.perf_counter() - t0
            for v in (s2, m2, xv):
                ir.dec_ref_count(v)
            if i >= 1:
                ts.append(dt)
        t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
        if world > 1:
            td.all_reduce(t, op=td.ReduceOp.MAX)
        return float(t.item())

    e2e_s = e2e_leg(lambda: ir.array_shard_local(T.F32, ptr=hp.value, n=n_local))
    e2e = {"value": total_bytes / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": N_TOTAL * 4,
           "d2h_bytes_per_step": 8 * world, "steps": e2e_steps, "seconds_per_step": e2e_s, "host_buffer": "pinned (vkjit_host_alloc)"}
    np_s = e2e_leg(lambda: ir.array_shard_local(T.F32, x_np))
    e2e_default = {"value": total_bytes / np_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": N_TOTAL * 4, "d2h_bytes_per_step": 8 * world,
                   "steps": e2e_steps, "seconds_per_step": np_s, "h2d_GBps_per_gpu": n_local * 4 / np_s / 1e9,
                   "host_buffer": "pageable NumPy array through the library's pinned staging ring (threaded host memcpy)"}
    api.call("host_free", hp)
    del x_np
    assert abs(out2[0] - gpu_sum) <= 1e-5 * abs(gpu_sum) and out2[1] == gpu_max

    # ---- N>1 extras (every rank takes part): weak scaling of the same reduction (2^28 lanes PER GPU) and
    # the sharded fused elementwise trace E28 (2^28 lanes in total, no collective at all)
    mgpu_extras = None
    if world > 1 and not args.no_extras:
        try:  # extras never take the headline down (an error here is the same on every rank)
            mgpu_extras = {}
            big = uniform_trace(ir, ir.arange_sharded(T.U32, N_TOTAL * world), SEED_R28)
            ir.eval([big])
            ts = []
            for i in range(3 + 10):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                r1 = ir.reduce(Red.Sum, big); r2 = ir.reduce(Red.Max, big)
                b.record(stream)
                vk.sync()
                ir.dec_ref_count(r1); ir.dec_ref_count(r2)
                if i >= 3:
                    ts.append(a.elapsed_time(b))
            t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            mgpu_extras["weak_R28_per_gpu"] = {"ms_per_step": float(t.item()), "GBps": 2 * N_TOTAL * world * 4 / (float(t.item()) * 1e-3) / 1e9,
                                               "lanes_per_gpu": N_TOTAL, "note": "sum+max, 2^28 lanes per GPU (weak scaling), fused P2P all-reduce" if args.collective == "p2p" else "NCCL"}
            ir.dec_ref_count(big)
            ex_x = uniform_trace(ir, lanes, 0xB2000011)
            ex_y = uniform_trace(ir, lanes, 0xB2000012)
            ir.eval([ex_x, ex_y])
            half = ir.const_f32(0.5)
            ts = []
            K = 8   # a shard's output (128 MiB at N = 8) is about the size of the L2: timed alone, part of its write-back
                    # falls behind the closing event.  K evals back to back into K different arrays: the dirty lines of
                    # eval k are written back under eval k + 1, only the last tail (<= 1/K of a third of the traffic) is missed
            for i in range(3 + 6):
                flush_l2()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                zs = []
                for _ in range(K):
                    z = ir.add(ir.mul(ex_x, ex_y), half); ir.eval([z]); zs.append(z)
                b.record(stream)
                vk.sync()
                for z in zs:
                    ir.dec_ref_count(z)
                if i >= 3:
                    ts.append(a.elapsed_time(b) / K)
            t = torch.tensor([sum(ts) / len(ts)], device=dev, dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            mgpu_extras["E28_sharded_elementwise"] = {"ms": float(t.item()), "GBps": 12 * N_TOTAL / (float(t.item()) * 1e-3) / 1e9,
                                                       "note": "z = x*y + c over 2^28 lanes in total, contiguous shards, no collective; 8 evals back to back per event pair, "
                                                               "every output a different array (the write-back of one lands under the next)"}
            ir.dec_ref_count(ex_x); ir.dec_ref_count(ex_y)
        except Exception as ex:
            mgpu_extras = {"error": repr(ex)}

    line = None
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": dict(CONFIG),
            "setup": {"n_per_gpu": n_local,
                      "parallelism": (f"contiguous 1-D shards x{world}, per-GPU partial + " + ("all-reduce fused into the reduce kernel's last CTA over NVLink peer memory (P2P mailbox)" if args.collective == "p2p" else "NCCL all-reduce")) if world > 1 else "single GPU",
                      "l2": "the steps rotate over 4 independent arrays (4 GiB/N per GPU against 126 MB of L2), each re-read only after 3 GiB/N of other reads; no flush inside the timed region",
                      "timing": "K steps back to back between one pair of CUDA events on the backend stream; barrier + synchronize on both sides, "
                                "plus (N>1) an untimed in-stream device barrier right before the first event; max over ranks"},
            "roofline": roofline, "e2e": e2e, "e2e_default_numpy": e2e_default, "gpu_launches": launches, "clocks": clocks, "isolated": isolated,
            "wall_s_timed_region": t_wall, "host_issue_us_per_reduction": t_issue / (2 * args.steps) * 1e6,
            "result": {"sum": gpu_sum, "max": gpu_max},
            "hbm_frac_whole_job": value / (peak * world),
        }
        if not args.no_cpu_baseline:
            # rank 0 has the host cores: the CPU arm runs (and checks the GPU result) at every N
            # (cpu_baseline is REPORTED at N=1 only — at N>1 the other ranks' processes share the host; there the child runs one pass for the check)
            cb = cpu_arm_subprocess() if world == 1 else cpu_arm_subprocess(steps=1, warmup=0)
            if world == 1:
                line["cpu_baseline"] = dict(cb["cpu_baseline"], steps=cb["steps"], warmup=cb["warmup"],
                                            how="child process `bench.py --impl reference` (the reference arm itself)")
            osum, omax = cb["result"]["sum"], cb["result"]["max"]
            tol = 1e-6 * LOG2N * abs(osum)
            line["check"] = {"sum_abs_err": abs(gpu_sum - osum), "sum_tol": tol, "max_equal": gpu_max == omax,
                             "ok": bool(abs(gpu_sum - osum) <= tol and gpu_max == omax), "against": "oracle port (CPU), full 2^28 lanes"}
        if world > 1:
            line["extras"] = mgpu_extras
        if world == 1 and not args.no_extras:
            try:
                import bench_extras
                line["extras"] = bench_extras.run(ir, vk, stream, flush_l2, peak)
                if not args.no_cpu_baseline:   # CPU legs of the same configs (child process, bounded samples)
                    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "LOCAL_WORLD_SIZE")}
                    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--extras"], capture_output=True,
                                       text=True, timeout=900, env=env, cwd=ROOT)
                    cpu = [json.loads(ln) for ln in r.stdout.splitlines() if ln.startswith("{")]
                    if r.returncode == 0 and cpu:
                        for k, v in cpu[-1]["extras_cpu_baseline"].items():
                            if k in line["extras"]:
                                line["extras"][k]["cpu_baseline"] = v
                    else:
                        line["extras"]["cpu_baseline_error"] = r.stderr[-400:]
                lc = line["extras"].get("LAUNCH_cached_trace", {})
                line["cached_trace_launch_us"] = lc.get("vkjit_eval_us_median")   # second half of BASELINE.json's metric
            except Exception as ex:  # extras never take the headline down
                line["extras"] = {"error": repr(ex)}
    barrier()
    ir.close()
    if world > 1:
        api.call("dist_shutdown")
        td.destroy_process_group()
    if rank == 0:
        if world > 1 and not args.no_extras:
            try:
                line["mgpu_parity"] = mgpu_parity(world)
            except Exception as ex:  # never takes the headline down
                line["mgpu_parity"] = {"ok": False, "error": repr(ex)}
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--collective", default="p2p", choices=["p2p", "nccl"], help="N>1: fused peer-memory all-reduce (default) or NCCL")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--extras", action="store_true", help="with --impl reference: the CPU legs of the secondary configs instead of the headline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference" and args.extras:
        cpu_extras_arm()
    elif args.impl == "reference":
        reference_arm(args)
    else:
        ours(args)


if __name__ == "__main__":
    main()
