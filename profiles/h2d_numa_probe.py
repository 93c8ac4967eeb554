"""Concurrent H2D bandwidth of N ranks (one per GPU) from pinned host memory, with and without binding each rank to the
CPUs its GPU is attached to (nvidia-smi topo: CPU affinity) BEFORE the pinned buffer is allocated (first touch decides
the NUMA node).  Explains the e2e figure at N = 8 (23 GB/s per GPU against 55 GB/s at N = 1).
    python profiles/h2d_numa_probe.py N"""
import os, subprocess, sys, time

def worker(rank, world, bind, t0):
    import torch
    if bind:
        out = subprocess.run(["nvidia-smi", "topo", "-C", "-i", str(rank)], capture_output=True, text=True).stdout
        # e.g. "CPU Affinity of GPU 0: 0-55,112-167" style output varies; fall back to parsing topo -m
        cpus = None
        for tok in out.replace(",", " ").split():
            pass
        m = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout.splitlines()
        hdr = [l for l in m if l.startswith("\tGPU0") or l.strip().startswith("GPU0")]
        for l in m:
            if l.startswith(f"GPU{rank}\t") or l.startswith(f"GPU{rank} "):
                cols = l.split("\t")
                for c in cols:
                    c = c.strip()
                    if c and all(ch.isdigit() or ch in "-," for ch in c) and ("-" in c or "," in c):
                        cpus = c; break
        if cpus:
            s = set()
            for part in cpus.split(","):
                a, _, b = part.partition("-")
                s.update(range(int(a), int(b or a) + 1))
            os.sched_setaffinity(0, s)
    torch.cuda.set_device(rank)
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True); h.fill_(1)
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True); torch.cuda.synchronize()
    while time.time() < t0: pass                      # crude common start
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(12): d.copy_(h, non_blocking=True)
    b.record(); torch.cuda.synchronize()
    print(f"rank {rank} bind={bind} affinity={len(os.sched_getaffinity(0))} cpus  H2D {12 * n / (a.elapsed_time(b) * 1e-3) / 1e9:.1f} GB/s", flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 2:
        worker(int(sys.argv[2]), int(sys.argv[1]), sys.argv[3] == "1", float(sys.argv[4]))
    else:
        world = int(sys.argv[1])
        print(subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout)
        print(subprocess.run("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'", shell=True, capture_output=True, text=True).stdout)
        for bind in ("0", "1"):
            for w in ([1, world] if world > 1 else [1]):
                t0 = time.time() + 25
                ps = [subprocess.Popen([sys.executable, __file__, str(w), str(r), bind, str(t0)]) for r in range(w)]
                for p in ps: p.wait()
                print(f"--- {w} rank(s), bind={bind} done", flush=True)
