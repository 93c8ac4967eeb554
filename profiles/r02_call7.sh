#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_c7_pytest.log 2>&1; tail -5 $O/r02_c7_pytest.log
python bench.py --steps 20 --warmup 5 > $O/r02_c7_bench_n1.json 2> $O/r02_c7_bench_n1.err; tail -3 $O/r02_c7_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c7_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'e2e_np', d['e2e_default_numpy'], d.get('cpu_baseline',{}).get('value'))
for k,v in d['extras'].items(): print(k, json.dumps(v)[:330])
PY
VKJIT_FAST_MATH=1 python - <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch, vkjit_b200 as vk, monte_carlo
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
fb = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
def flush():
    with torch.cuda.stream(stream): fb.sum()
r = monte_carlo.bench(vk, stream, flush)
print("M26 VKJIT_FAST_MATH=1", json.dumps({k: r[k] for k in ("kernel_ms", "Glanes_per_s", "compile_ms_cold")}))
PY
