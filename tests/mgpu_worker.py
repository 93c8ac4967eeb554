"""One rank of the multi-GPU parity test: one process per GPU.

    torchrun ... tests/mgpu_worker.py             rendezvous through torch.distributed (NCCL process group)
    RANK=r WORLD_SIZE=n LOCAL_RANK=r MASTER_ADDR=127.0.0.1 MASTER_PORT=p python tests/mgpu_worker.py --no-torch
                                                  no torch at all: vkjit_dist_init_env (native TCP rendezvous), what a
                                                  vkjit-rust or C caller uses
"""
import math
import os
import sys

import numpy as np

NO_TORCH = "--no-torch" in sys.argv
if not NO_TORCH:
    import torch
    import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import vkjit_b200 as vk  # noqa: E402
from oracle_lib import OracleIr  # noqa: E402
from vkjit_b200 import dist  # noqa: E402
from vkjit_b200.ir import Bop, Ir, Red, VarType as T  # noqa: E402


def trace_u32(ir, lanes):
    h = ir.mul(ir.bop(Bop.Xor, lanes, ir.const_u32(0x9E3779B9)), ir.const_u32(747796405))
    return ir.bop(Bop.Shr, h, ir.const_u32(7))


def trace_f32(ir, lanes):
    return ir.mul(ir.cast(ir.bop(Bop.Shr, trace_u32(ir, lanes), ir.const_u32(1)), T.F32), ir.const_f32(2.0 ** -24))


def main():
    local = int(os.environ["LOCAL_RANK"])
    if NO_TORCH:
        assert "torch" not in sys.modules
        rank, world = dist.init_from_env()      # binds the device, NCCL communicator, peer mailboxes
    else:
        torch.cuda.set_device(local)
        td.init_process_group("nccl", device_id=torch.device("cuda", local))
        vk.init(local)
        rank, world = dist.init_from_torch(torch.device("cuda", local))
    for p2p, n in [(m, k) for m in (True, False) for k in (5, 1000, (1 << 22) + 3)]:
        dist.set_p2p(p2p)  # fused reduce + all-reduce over NVLink peer memory, then the NCCL path
        ir = Ir()
        lanes = ir.arange_sharded(T.U32, n)
        lo, hi = dist.shard_range(n, rank, world)
        xu, xf = trace_u32(ir, lanes), trace_f32(ir, lanes)
        assert ir.is_sharded(xu)
        got = {}
        # every rank takes part in the collective, also when its shard of a tiny array is empty
        for name, v, ty in (("u", xu, T.U32), ("f", xf, T.F32)):
            for r in (Red.Sum, Red.Min, Red.Max):
                got[(name, r)] = ir.as_slice(ir.reduce(r, v), ty)[0]
        o = OracleIr()
        ol = o.arange(T.U32, n)
        ou, of = trace_u32(o, ol), trace_f32(o, ol)
        o.eval([ou, of])
        if hi > lo:
            # elementwise: this rank's lanes equal the global lanes [lo, hi) bit for bit, no collective
            assert ir.as_slice_eval(xu, T.U32).tobytes() == o.as_slice(ou, T.U32)[lo:hi].tobytes()
            assert ir.as_slice_eval(xf, T.F32).tobytes() == o.as_slice(of, T.F32)[lo:hi].tobytes()
        if True:
            for r in (Red.Sum, Red.Min, Red.Max):
                e = o.as_slice(o.reduce(r, ou), T.U32)[0]
                assert got[("u", r)] == e, (n, r, got[("u", r)], e)       # bit-exact, replicated on every rank
            es = float(o.as_slice(o.reduce(Red.Sum, of), T.F32)[0])
            assert abs(float(got[("f", Red.Sum)]) - es) <= 1e-6 * max(1.0, math.log2(n)) * abs(es)
            for r in (Red.Min, Red.Max):
                assert got[("f", r)] == o.as_slice(o.reduce(r, of), T.F32)[0]
        # sharded prefix sum: every rank's result is its slice of the scan over the GLOBAL range
        for excl in (True, False):
            ps = ir.prefix_sum(xu, excl)
            assert ir.is_sharded(ps) and ir.size(ps) == hi - lo
            want = o.as_slice(o.prefix_sum(ou, excl), T.U32)[lo:hi]
            if hi > lo:
                assert ir.as_slice(ps, T.U32).tobytes() == want.tobytes(), (n, p2p, excl)
        # the same on an UNEVALUATED sharded trace: local total from the fused trace -> reduce kernel, then the
        # fused trace -> scan kernel seeded with the lower ranks' totals; the operand is never materialised
        xu2 = ir.add(trace_u32(ir, lanes), ir.const_u32(7))
        ps = ir.prefix_sum(xu2, True)
        assert not ir.is_buffer(xu2) and ir.is_sharded(ps) and ir.size(ps) == hi - lo
        want = o.as_slice(o.prefix_sum(o.add(ou, o.const_u32(7)), True), T.U32)[lo:hi]
        if hi > lo:
            assert ir.as_slice(ps, T.U32).tobytes() == want.tobytes(), (n, p2p, "fused")
        # sharded compress: every rank compacts its shard; results are ragged shards of the GLOBAL result
        for fused in (True, False):
            m_d = ir.neq(ir.bop(Bop.And, trace_u32(ir, lanes), ir.const_u32(4)), ir.const_u32(0))
            v_d = ir.add(trace_u32(ir, lanes), ir.const_u32(1))
            if not fused and hi > lo:
                ir.eval([m_d, v_d])                                 # hand-written kernel on bound arrays
            m_o = o.neq(o.bop(Bop.And, ou, o.const_u32(4)), o.const_u32(0))
            want_idx, cg = o.compress(m_o)
            want_val, _ = o.compress_values(o.add(ou, o.const_u32(1)), m_o)
            want_idx, want_val = o.as_slice(want_idx, T.U32), o.as_slice(want_val, T.U32)
            gi, c1 = ir.compress(m_d)
            gv, c2 = ir.compress_values(v_d, m_d)
            assert c1 == c2 == cg, (n, p2p, fused, c1, c2, cg)         # global count, replicated
            assert ir.is_sharded(gi) and ir.is_sharded(gv)
            off, k = ir.shard_base(gi), ir.size(gi)
            assert ir.shard_base(gv) == off and ir.size(gv) == k
            # this rank's ragged shard = the selected lanes of [lo, hi): its offset and size follow from the oracle's mask
            mo = o.as_slice_eval(m_o, T.Bool) != 0
            assert off == int(mo[:lo].sum()) and k == int(mo[lo:hi].sum()), (n, p2p, fused, off, k)
            if k:
                assert ir.as_slice(gi, T.U32).tobytes() == want_idx[off:off + k].tobytes(), (n, p2p, fused)   # GLOBAL lane numbers
                assert ir.as_slice(gv, T.U32).tobytes() == want_val[off:off + k].tobytes(), (n, p2p, fused)
        ir.close(); o.close()
    # Mailbox flow control: far more than kMailSlots = 64 exchanges back to back with no host sync, one
    # rank lagging on the host — all-reduces (fused into the reduce kernel and stand-alone) and exclusive scans over
    # ranks mixed.  Every exchange waits for all ranks, so a fast rank can never lap the 64-slot ring.
    import time
    dist.set_p2p(True)
    ir = Ir()
    n = 40000 * world + 12
    lo, hi = dist.shard_range(n, rank, world)
    lanes = ir.arange_sharded(T.U32, n)
    xb = ir.add(lanes, ir.const_u32(0)); ir.eval([xb])                      # a bound sharded array: hand-written reduce
    sums, scans = [], []
    for i in range(300):
        if rank == world - 1 and i % 75 == 10:
            time.sleep(0.05)                                                # this rank falls 50 ms behind
        if i % 3 == 0:
            sums.append((i, ir.reduce(Red.Sum, ir.add(lanes, ir.const_u32(i)))))   # fused trace -> reduce + stand-alone exchange
        elif i % 3 == 1:
            sums.append((0, ir.reduce(Red.Sum, xb)))                        # exchange inside the reduce kernel's last CTA
        else:
            scans.append(ir.prefix_sum(xb, True))                           # exclusive scan over ranks of the shard totals
    M = 1 << 32
    for i, v in sums:
        assert int(ir.as_slice(v, T.U32)[0]) == (n * (n - 1) // 2 + n * i) % M, i
    for v in scans:
        got = ir.as_slice(v, T.U32)
        if hi > lo:
            assert int(got[0]) == (lo * (lo - 1) // 2) % M and int(got[-1]) == ((hi - 1) * (hi - 2) // 2) % M
    ir.close()
    st = vk.stats()
    assert st["collectives"] > 0 or world == 1
    if NO_TORCH:
        assert "torch" not in sys.modules
        ir = Ir()                                # a last sharded reduction doubles as the barrier before shutdown
        ir.as_slice(ir.reduce(Red.Sum, ir.arange_sharded(T.U32, 64 * world)), T.U32)
        ir.close()
        dist.shutdown()
    else:
        td.barrier()
        dist.shutdown()
        td.destroy_process_group()
    print(f"rank {rank}/{world} ok collectives={st['collectives']}")


if __name__ == "__main__":
    main()
