"""Multi-GPU plumbing for the sharded paths (one process per GPU).

`torch.distributed` is only the rendezvous: rank 0 asks the native library for an NCCL unique id,
the id travels through one broadcast, and every rank hands it to `vkjit_dist_init`.  From then on
sharded reductions combine their per-GPU partials inside the native library (NCCL all-reduce on
the backend stream).  Elementwise traces shard with no collective at all.
"""
from __future__ import annotations

import ctypes as C

from ._capi import product_api

UNIQUE_ID_BYTES = 128


def shard_range(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of a global 1-D range: multiples of 4 lanes (16-byte aligned
    128-bit accesses), the last rank takes the ragged tail.  Same rule as the native library."""
    lo, hi = C.c_size_t(), C.c_size_t()
    product_api().call("shard_range", n, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def exchange_unique_id(make_id, rank: int, device=None) -> bytes:
    """Broadcast a 128-byte id from rank 0 over the default torch.distributed group.
    Works with any backend (gloo on CPU in the tests, nccl on the GPUs)."""
    import torch
    import torch.distributed as td
    buf = torch.zeros(UNIQUE_ID_BYTES, dtype=torch.uint8, device=device or "cpu")
    if rank == 0:
        raw = make_id()
        assert len(raw) == UNIQUE_ID_BYTES
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    td.broadcast(buf, 0)
    return bytes(buf.cpu().tolist())


def native_unique_id() -> bytes:
    host = (C.c_uint8 * UNIQUE_ID_BYTES)()
    product_api().call("dist_unique_id", host)
    return bytes(host)


def init_from_torch(device=None, p2p: bool = True):
    """Call after torch.distributed.init_process_group and vkjit_b200.init(local_rank)."""
    import torch.distributed as td
    rank, world = td.get_rank(), td.get_world_size()
    raw = exchange_unique_id(native_unique_id, rank, device) if world > 1 else bytes(UNIQUE_ID_BYTES)
    product_api().call("dist_init", rank, world, C.create_string_buffer(raw, UNIQUE_ID_BYTES))
    if world > 1 and p2p:
        open_mailboxes(rank, world, device)
    return rank, world


def open_mailboxes(rank: int, world: int, device=None):
    """Exchange the cudaIpc handles of the per-GPU mailboxes (64 bytes each) and map them, which
    switches sharded reductions to the fused reduce + all-reduce kernel over NVLink peer memory."""
    import torch
    import torch.distributed as td
    api = product_api()
    mine = (C.c_uint8 * 64)()
    api.call("dist_mailbox_handle", mine)
    t = torch.tensor(list(bytes(mine)), dtype=torch.uint8, device=device or "cpu")
    allh = [torch.zeros(64, dtype=torch.uint8, device=device or "cpu") for _ in range(world)]
    td.all_gather(allh, t)
    raw = b"".join(bytes(h.cpu().tolist()) for h in allh)
    api.call("dist_mailbox_open", C.create_string_buffer(raw, 64 * world), world)


def init_from_env():
    """Multi-GPU bring-up WITHOUT torch (native TCP rendezvous, csrc/rendezvous.cpp): RANK / WORLD_SIZE / LOCAL_RANK /
    MASTER_ADDR / MASTER_PORT from the environment, as any launcher sets them.  Binds the device (LOCAL_RANK), creates
    the NCCL communicator and maps the peer mailboxes.  Returns (rank, world)."""
    api = product_api()
    api.call("dist_init_env")
    r, w = C.c_int32(), C.c_int32()
    api.call("dist_info", C.byref(r), C.byref(w))
    return r.value, w.value


def rendezvous(rank: int, world: int, blob64: bytes, root128: bytes, addr: str = "127.0.0.1", port: int = 29501, timeout_s: float = 60.0):
    """The exchange alone (needs no device): every rank's 64-byte blob to everyone, rank 0's 128-byte blob to everyone."""
    assert len(blob64) == 64 and len(root128) == 128
    root = C.create_string_buffer(root128, 128)
    out = C.create_string_buffer(64 * world)
    product_api().call("debug_rendezvous", rank, world, addr.encode(), port, C.create_string_buffer(blob64, 64), root, out, timeout_s)
    return bytes(root.raw), [bytes(out.raw[64 * r:64 * (r + 1)]) for r in range(world)]


def shutdown():
    product_api().call("dist_shutdown")


def set_p2p(on: bool):
    """Fused mailbox all-reduce (True) or NCCL (False); every rank must make the same call."""
    product_api().call("dist_set_p2p", 1 if on else 0)
