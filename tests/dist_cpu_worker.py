"""world_size-2 gloo worker (CPU): the host-side logic of the sharded reduce path.
Each rank evaluates ITS shard with the oracle, the partials are combined with an all-reduce, and
the result must equal the oracle's evaluation of the whole range."""
import os
import sys

import numpy as np
import torch
import torch.distributed as td

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle_lib import oracle_api  # noqa: E402
from vkjit_b200 import dist  # noqa: E402
from vkjit_b200.ir import Bop, Ir, Red, VarType as T  # noqa: E402


def trace(ir, lanes):
    h = ir.mul(ir.bop(Bop.Xor, lanes, ir.const_u32(0x9E3779B9)), ir.const_u32(747796405))
    return ir.bop(Bop.Shr, h, ir.const_u32(7))


def main():
    td.init_process_group("gloo")
    rank, world = td.get_rank(), td.get_world_size()
    api = oracle_api()
    n = 100003

    # 1) unique-id plumbing: every rank ends up with rank 0's 128 bytes
    raw = dist.exchange_unique_id(lambda: bytes((i * 7 + 3) % 256 for i in range(128)), rank)
    assert raw == bytes((i * 7 + 3) % 256 for i in range(128))

    # 2) shards tile the range, 16-byte aligned
    lo, hi = dist.shard_range(n, rank, world)
    spans = [None] * world
    td.all_gather_object(spans, (lo, hi))
    assert spans[0][0] == 0 and spans[-1][1] == n
    for a, b in zip(spans, spans[1:]):
        assert a[1] == b[0] and (a[1] - a[0]) % 4 == 0

    # 3) sharded elementwise trace + reduction: local partial, then all-reduce (sum/min/max)
    import ctypes as C
    ir = Ir(_api=api)
    out = C.c_uint32()
    api.call("arange_shard", ir._h, T.U32, n, rank, world, C.byref(out))
    x = trace(ir, out.value)
    part = {r: int(ir.as_slice(ir.reduce(r, x), T.U32)[0]) for r in (Red.Sum, Red.Min, Red.Max)}
    ir.eval([x])
    assert ir.size(x) == hi - lo
    ts = torch.tensor([part[Red.Sum]], dtype=torch.int64); td.all_reduce(ts, op=td.ReduceOp.SUM)
    tmin = torch.tensor([part[Red.Min]], dtype=torch.int64); td.all_reduce(tmin, op=td.ReduceOp.MIN)
    tmax = torch.tensor([part[Red.Max]], dtype=torch.int64); td.all_reduce(tmax, op=td.ReduceOp.MAX)
    full = Ir(_api=api)
    xf = trace(full, full.arange(T.U32, n))
    exp = {r: int(full.as_slice(full.reduce(r, xf), T.U32)[0]) for r in (Red.Sum, Red.Min, Red.Max)}
    assert int(ts.item()) % (1 << 32) == exp[Red.Sum], (int(ts.item()), exp)
    assert int(tmin.item()) == exp[Red.Min] and int(tmax.item()) == exp[Red.Max]
    # the shard holds exactly the global lanes [lo, hi)
    full.eval([xf])
    assert np.array_equal(ir.as_slice(x, T.U32), full.as_slice(xf, T.U32)[lo:hi])
    # 4) sharded compress: per-shard compaction, global lane numbers = shard base + local index, ragged result
    #    placed by an exclusive scan of the per-rank counts — the composition the device path implements
    m = ir.neq(ir.bop(Bop.And, x, ir.const_u32(4)), ir.const_u32(0))
    idx_local, k = ir.compress(m)
    val_local, _ = ir.compress_values(x, m)
    counts = [None] * world
    td.all_gather_object(counts, k)
    off = sum(counts[:rank])
    mf = full.neq(full.bop(Bop.And, xf, full.const_u32(4)), full.const_u32(0))
    idx_full, kf = full.compress(mf)
    val_full, _ = full.compress_values(xf, mf)
    assert sum(counts) == kf
    assert np.array_equal(ir.as_slice(idx_local, T.U32) + np.uint32(lo), full.as_slice(idx_full, T.U32)[off:off + k])
    assert np.array_equal(ir.as_slice(val_local, T.U32), full.as_slice(val_full, T.U32)[off:off + k])
    td.barrier()
    td.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
