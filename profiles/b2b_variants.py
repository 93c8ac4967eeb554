"""Back-to-back sharded reductions (torchrun, N GPUs): which ingredient of bench.py's loop costs what.
Variants: same array / 4 rotating arrays, sum only / sum+max alternating, results dropped at once / two steps later.
    torchrun --nproc-per-node N profiles/b2b_variants.py [log2 lanes per GPU]"""
import os, sys, time
import torch, torch.distributed as td
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
from bench import uniform_trace
from vkjit_b200 import dist
from vkjit_b200.ir import Ir, Red, VarType as T
local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
td.init_process_group("nccl", device_id=dev)
vk.init(local)
rank, world = dist.init_from_torch(dev)
stream = torch.cuda.ExternalStream(vk.stream_ptr(), device=dev)
ir = Ir()
log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 27
n = (1 << log2n) * world
lanes = ir.arange_sharded(T.U32, n)
xs = [uniform_trace(ir, lanes, 0xB2000021 + i) for i in range(4)]
for x in xs:
    ir.eval([x])
vk.sync()
reps = 200
for rot in (1, 4):
    for ops in ((Red.Sum,), (Red.Sum, Red.Max)):
        for lag in (0, 4):
            pend = []
            def one(j):
                pend.append(ir.reduce(ops[j % len(ops)], xs[j % rot]))
                while len(pend) > lag:
                    ir.dec_ref_count(pend.pop(0))
            for j in range(20):
                one(j)
            vk.sync(); td.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record(stream)
            for j in range(reps):
                one(j)
            b.record(stream)
            host_us = (time.perf_counter() - t0) * 1e6 / reps
            vk.sync()
            while pend:
                ir.dec_ref_count(pend.pop(0))
            t = torch.tensor([a.elapsed_time(b) * 1e3 / reps, host_us], device=dev, dtype=torch.float64)
            td.all_reduce(t, op=td.ReduceOp.MAX)
            if rank == 0:
                print(f"world={world} lanes/gpu=2^{log2n} arrays={rot} ops={len(ops)} drop_lag={lag}: {t[0].item():.2f} us per reduce, host enqueue {t[1].item():.2f} us", flush=True)
td.barrier(); dist.shutdown(); td.destroy_process_group()
