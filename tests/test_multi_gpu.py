"""Sharded elementwise + reduce across the GPUs of one box (one process per GPU, NCCL all-reduce of
the per-GPU partials).  Skipped on a single-GPU box; run with `gpurun --gpus 2 -- pytest -m gpu`."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_reduce_matches_oracle_on_all_gpus():
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else min(ngpu, 8)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    for k in range(world):
        assert f"rank {k}/{world} ok" in r.stdout


def test_sharded_paths_without_torch():
    """The same worker started as plain processes (no torchrun, no torch import): vkjit_dist_init_env's native TCP
    rendezvous brings up NCCL and the peer mailboxes — the route a vkjit-rust / C caller takes (VERDICT r01 task 7)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if ngpu < 4 else min(ngpu, 8)
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT="29571")
        procs.append(subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py"), "--no-torch"], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True))
    outs = [p.communicate(timeout=600) for p in procs]
    for r, (p, (o, e)) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, (r, o[-2000:], e[-4000:])
        assert f"rank {r}/{world} ok" in o
