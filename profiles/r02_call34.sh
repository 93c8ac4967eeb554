#!/bin/bash
# N = 8: bench (driver launch line) with the final code of the round
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 bash profiles/bench_n.sh 8 r02final --steps 20 --warmup 5
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n8_r02final.json'))
print('check', d.get('check'), 'mgpu_parity', json.dumps(d.get('mgpu_parity'))[:300], 'clocks', d.get('clocks'))
PY
timeout 300 bash profiles/bench_n.sh 4 r02final --steps 20 --warmup 5 --no-extras
