"""N>1 host logic on CPU: world_size-2 gloo run of tests/dist_cpu_worker.py."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gloo_world2_sharded_reduce_logic():
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "dist_cpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "rank 0 ok" in r.stdout and "rank 1 ok" in r.stdout


def test_native_rendezvous_three_ranks_no_torch(tmp_path):
    """csrc/rendezvous.cpp (what vkjit_dist_init_env uses instead of torch.distributed): three plain processes, the
    later ones started BEFORE rank 0 listens; every rank must end up with rank 0's 128-byte blob and all three 64-byte
    blobs in rank order; a stray connection to the port is ignored."""
    script = tmp_path / "rdzv.py"
    script.write_text('''
import sys, time
sys.path.insert(0, %r)
from vkjit_b200 import dist
rank, world, port = int(sys.argv[1]), 3, int(sys.argv[2])
if rank == 0:
    time.sleep(0.7)          # the other ranks are already retrying
mine = bytes([rank + 1]) * 64
root = (bytes(range(128)) if rank == 0 else bytes(128))
root, blobs = dist.rendezvous(rank, world, mine, root, port=port, timeout_s=30)
assert root == bytes(range(128)), rank
assert blobs == [bytes([r + 1]) * 64 for r in range(world)], rank
print("rank", rank, "ok")
''' % ROOT)
    port = 29541 + (os.getpid() % 200)
    procs = [subprocess.Popen([sys.executable, str(script), str(r), str(port)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for r in (2, 1, 0)]
    import socket
    import time
    time.sleep(1.0)
    try:                      # a stray client that says nothing sensible
        s = socket.create_connection(("127.0.0.1", port), timeout=1); s.send(b"hello"); s.close()
    except OSError:
        pass
    outs = [p.communicate(timeout=90) for p in procs]
    for p, (o, e) in zip(procs, outs):
        assert p.returncode == 0, o + e
    assert sorted(o.strip() for o, _ in outs) == ["rank 0 ok", "rank 1 ok", "rank 2 ok"]
