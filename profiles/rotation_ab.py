"""How should a stream of reductions be timed?  Single GPU, shard sizes of the strong-scaling run (2^28 / N lanes).
  isolated : L2 evicted by a 256 MiB read, one CUDA-event pair per reduction (bench.py's method up to now)
  rotate R : R independent input arrays used round-robin, sum/max alternating, ONE event pair around 2K reductions
If rotating over 4 arrays still left lines in L2, 8 arrays would be slower than 4 — they are not.
    python profiles/rotation_ab.py > gpurun_out/rotation_ab.json"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vkjit_b200 as vk  # noqa: E402
from bench import uniform_trace  # noqa: E402
from vkjit_b200.ir import Ir, Red, VarType as T  # noqa: E402

vk.init(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(vk.stream_ptr(), device=dev)
ir = Ir()
flush_buf = torch.zeros(64 << 20, dtype=torch.float32, device=dev)
out = {}
K = 50
for log2n in (28, 27, 26, 25):
    n = 1 << log2n
    xs = [uniform_trace(ir, ir.arange(T.U32, n), 0xB2000021 + i) for i in range(8)]
    for x in xs:
        ir.eval([x])
    vk.sync()
    res = {}
    ts = []
    for i in range(3 + 2 * K):
        with torch.cuda.stream(stream):
            flush_buf.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        r = ir.reduce(Red.Sum if i % 2 == 0 else Red.Max, xs[0])
        b.record(stream)
        vk.sync()
        ir.dec_ref_count(r)
        if i >= 3:
            ts.append(a.elapsed_time(b))
    res["isolated_us"] = 1e3 * sum(ts) / len(ts)
    for R in (1, 2, 4, 8):
        prev = None
        for rep in range(2):   # first pass = warm-up
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            vk.sync()
            a.record(stream)
            for j in range(2 * K):
                r = ir.reduce(Red.Sum if j % 2 == 0 else Red.Max, xs[j % R])
                if prev is not None:
                    ir.dec_ref_count(prev)
                prev = r
            b.record(stream)
            vk.sync()
        ir.dec_ref_count(prev)
        res[f"rotate{R}_us"] = 1e3 * a.elapsed_time(b) / (2 * K)
    res["stream_floor_us_at_peak"] = n * 4 / 6551e9 * 1e6
    out[f"2^{log2n}"] = res
    for x in xs:
        ir.dec_ref_count(x)
print(json.dumps(out, indent=1))
