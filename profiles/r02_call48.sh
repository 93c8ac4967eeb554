#!/bin/bash
# full GPU test tier + N=1 bench with the driver flags (final code of the round, after the second-session kernels)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > $O/r02_c48_pytest.log 2>&1; tail -4 $O/r02_c48_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c48_smoke.log 2>&1; tail -2 $O/r02_c48_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_c48_bench_n1.json 2> $O/r02_c48_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_c48_bench_ref.json 2> $O/r02_c48_bench_ref.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c48_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'np', d['e2e_default_numpy']['value'], d.get('cpu_baseline',{}).get('value'), d['check']['ok'])
for k,v in d['extras'].items(): print(k, json.dumps(v)[:300])
r=json.loads(open('gpurun_out/r02_c48_bench_ref.json').read().strip().splitlines()[-1]); print('REF', r['value'])
PY
