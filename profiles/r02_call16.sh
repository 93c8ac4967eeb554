#!/bin/bash
# range-proven vk_math fast paths (M26) + look-back window spanning the whole co-resident grid (fused scans, 2 x 512)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --timeout 300 > $O/r02_c16_pytest.log 2>&1; tail -4 $O/r02_c16_pytest.log
{
echo "== default (look_wide 10 for 2 x 512)"; timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== VKJIT_LOOK_WIDE=5 (round-1/2 window)"; VKJIT_LOOK_WIDE=5 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== non-lag T=512 (look_wide 10)"; VKJIT_SCAN_T_NOLAG=512 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== everything lagged at 2 x 512 (look_wide 10)"; VKJIT_LAG_MAX_NODES=64 VKJIT_SCAN_T=512 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== VKJIT_LOOK_WIDE=12"; VKJIT_LOOK_WIDE=12 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
} | tee $O/r02_c16_fused_scan.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_c16_bench_n1.json 2> $O/r02_c16_bench_n1.err
VKJIT_NO_RANGES=1 timeout 300 python - > $O/r02_c16_m26_noranges.txt 2>&1 <<'PY'
import sys, json
sys.path.insert(0, '.')
import torch
import vkjit_b200 as vk
import monte_carlo
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
fb = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
def flush():
    with torch.cuda.stream(stream): fb.sum()
print(json.dumps(monte_carlo.bench(vk, stream, flush)))
PY
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c16_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], d['check']['ok'])
for k,v in d['extras'].items():
    if k.startswith(('C28','M26')): print(k, json.dumps(v)[:330])
print('NO_RANGES', open('gpurun_out/r02_c16_m26_noranges.txt').read()[-700:])
PY
