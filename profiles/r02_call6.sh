#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02_c6_pytest.log 2>&1; tail -5 $O/r02_c6_pytest.log
python profiles/fused_scan_ab.py > $O/r02_c6_fused_scan.json 2> $O/r02_c6_fused_scan.err; cat $O/r02_c6_fused_scan.json; tail -3 $O/r02_c6_fused_scan.err
python bench.py --steps 20 --warmup 5 > $O/r02_c6_bench_n1.json 2> $O/r02_c6_bench_n1.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c6_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], d['e2e']['value'], d.get('cpu_baseline'))
for k,v in d['extras'].items(): print(k, json.dumps(v)[:260])
PY
