#!/bin/bash
# Round 2, GPU call 2 (N GPUs of one box): sharded-reduction timeline, bench at the driver's flags and at --steps 100, multi-GPU parity tests.
cd "$GRAFT_REPO_ROOT" || exit 1
N=${1:-2}
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
VKJIT_REDUCE_TRACE=1 $TR --master-port 29611 profiles/reduce_timeline.py --steps 20 --out $O/r02_timeline_n${N}_aligned > $O/r02_timeline_n${N}_aligned.json 2> $O/r02_timeline_n${N}_aligned.err
VKJIT_REDUCE_TRACE=1 $TR --master-port 29612 profiles/reduce_timeline.py --steps 20 --no-align --out $O/r02_timeline_n${N}_noalign > $O/r02_timeline_n${N}_noalign.json 2> $O/r02_timeline_n${N}_noalign.err
$TR --master-port 29613 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $O/r02_bench_n${N}_s20.json 2> $O/r02_bench_n${N}_s20.err
$TR --master-port 29614 bench.py --gpus $N --steps 100 --warmup 5 --no-extras > $O/r02_bench_n${N}_s100.json 2> $O/r02_bench_n${N}_s100.err
$TR --master-port 29615 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > $O/r02_bench_n${N}_s20b.json 2> $O/r02_bench_n${N}_s20b.err
if [ "$2" = "tests" ]; then python -m pytest tests/test_multi_gpu.py -m gpu -x -q > $O/r02_pytest_mgpu_n${N}.log 2>&1; tail -3 $O/r02_pytest_mgpu_n${N}.log; fi
for f in $O/r02_bench_n${N}_s20.json $O/r02_bench_n${N}_s100.json $O/r02_bench_n${N}_s20b.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], d["value"], d["ms_per_step"], d.get("host_issue_us_per_reduction"), d["isolated"]["value"], d.get("check"))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
tail -c 600 $O/r02_timeline_n${N}_aligned.err
