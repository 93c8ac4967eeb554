"""Latency of the fused reduce+all-reduce vs NCCL on tiny sharded arrays (torchrun, N GPUs): isolates the
collective's own cost from streaming time and from rank skew."""
import os, sys, time
import torch, torch.distributed as td
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import vkjit_b200 as vk
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
for n in (1 << 12, 1 << 25):
    x = ir.cast(ir.arange_sharded(T.U32, n * world), T.F32)
    ir.eval([x])
    for mode in ("p2p", "nccl", "local"):
        if mode != "local":
            dist.set_p2p(mode == "p2p")
        xs = x
        if mode == "local":
            xs = ir.cast(ir.arange(T.U32, n), T.F32); ir.eval([xs])
        td.barrier(); vk.sync()
        reps = 200
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            ir.dec_ref_count(ir.reduce(Red.Sum, xs))
        vk.sync(); td.barrier()
        a.record(stream)
        for _ in range(reps):
            ir.dec_ref_count(ir.reduce(Red.Sum, xs))
        b.record(stream); vk.sync()
        t = torch.tensor([a.elapsed_time(b) * 1e3 / reps], device=dev, dtype=torch.float64)
        td.all_reduce(t, op=td.ReduceOp.MAX)
        if rank == 0:
            print(f"world={world} lanes/gpu={n} mode={mode}: {t.item():.2f} us per reduce (back-to-back, max over ranks)")
td.barrier(); dist.shutdown(); td.destroy_process_group()
