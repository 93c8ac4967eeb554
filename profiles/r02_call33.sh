#!/bin/bash
# N = 2: multi-GPU test tier + bench (driver launch line) + reference arm under torchrun
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 800 > $O/r02_c33_pytest_mgpu.log 2>&1; tail -4 $O/r02_c33_pytest_mgpu.log
timeout 600 bash profiles/bench_n.sh 2 r02final --steps 20 --warmup 5
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 2> $O/r02_c33_ref_n2.err | grep '^{' > $O/r02_c33_ref_n2.json; cut -c1-300 $O/r02_c33_ref_n2.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_r02final.json'))
print('check', d.get('check'), 'mgpu_parity', json.dumps(d.get('mgpu_parity'))[:300], 'clocks', d.get('clocks'))
PY
