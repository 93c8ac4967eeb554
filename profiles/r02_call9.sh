#!/bin/bash
# multi-GPU verification: parity tests (torchrun and torch-free), bench at the driver's flags (with the in-run parity leg), timeline
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$2" = "tests" ]; then timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 500 > $O/r02_c9_pytest_mgpu_n${N}.log 2>&1; tail -4 $O/r02_c9_pytest_mgpu_n${N}.log; fi
timeout 600 $TR --master-port 29713 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_c9_bench_n${N}.json 2> $O/r02_c9_bench_n${N}.err
[ "$3" = "lean" ] || timeout 300 $TR --master-port 29723 bench.py --gpus $N --steps 100 --warmup 5 --no-extras --no-cpu-baseline > $O/r02_c9_bench_n${N}_s100.json 2> $O/r02_c9_bench_n${N}_s100.err
VKJIT_REDUCE_TRACE=1 timeout 300 $TR --master-port 29733 profiles/reduce_timeline.py --steps 20 --out $O/r02_c9_timeline_n${N} > $O/r02_c9_timeline_n${N}.json 2> $O/r02_c9_timeline_n${N}.err
[ "$3" = "lean" ] || VKJIT_REDUCE_SPARE_CTAS=0 VKJIT_REDUCE_TRACE=1 timeout 300 $TR --master-port 29743 profiles/reduce_timeline.py --steps 20 --out $O/r02_c9_timeline_n${N}_spare0 > $O/r02_c9_timeline_n${N}_spare0.json 2> $O/r02_c9_timeline_n${N}_spare0.err
for f in $O/r02_c9_bench_n${N}.json $O/r02_c9_bench_n${N}_s100.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d.get("host_issue_us_per_reduction"), "iso", d["isolated"]["value"], d.get("check"), "e2e", d["e2e"]["value"], d.get("e2e_default_numpy", {}).get("value"), "parity", d.get("mgpu_parity"))
    print(json.dumps(d.get("extras"))[:900])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -c 400 $O/r02_c9_bench_n${N}.err
