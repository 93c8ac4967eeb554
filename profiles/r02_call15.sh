#!/bin/bash
# final N=1 record: GPU tests, smoke, bench at the driver's flags, reference arm, ncu launch list of the bench loop
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --timeout 300 > $O/r02_c15_pytest.log 2>&1; tail -4 $O/r02_c15_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02_c15_smoke.log 2>&1; tail -2 $O/r02_c15_smoke.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r02_c15_clocks.csv &
SMI=$!
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_c15_bench_n1.json 2> $O/r02_c15_bench_n1.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_c15_bench_ref.json 2> $O/r02_c15_bench_ref.err
kill $SMI
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_c15_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_c15_ncu_bench.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_kernel -s 6 -c 2 -o $O/r02_c15_reduce \
    python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_c15_ncu_reduce.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_c15_bench_n1.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline']['frac'], 'e2e', d['e2e']['value'], 'np', d['e2e_default_numpy']['value'], d.get('cpu_baseline',{}).get('value'), d['check'])
for k,v in d['extras'].items(): print(k, json.dumps(v)[:420])
r=json.loads(open('gpurun_out/r02_c15_bench_ref.json').read().strip().splitlines()[-1]); print('REF', r['value'])
PY
