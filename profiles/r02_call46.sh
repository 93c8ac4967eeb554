#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 bash profiles/bench_n.sh 2 r02final2 --steps 20 --warmup 5
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n2_r02final2.json'))
print('check', d.get('check',{}).get('ok'), 'mgpu_parity', d.get('mgpu_parity',{}).get('ok'), json.dumps(d.get('extras'))[:700])
PY
