#!/bin/bash
# bench.py on N GPUs (torchrun), JSON line to gpurun_out/bench_n${N}_$2.json, key figures on stdout.
N=$1; TAG=${2:-run}; shift; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N "$@" 2> gpurun_out/bench_n${N}_${TAG}.err | grep '^{' > gpurun_out/bench_n${N}_${TAG}.json
python - gpurun_out/bench_n${N}_${TAG}.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("n_gpus", "value", "ms_per_step", "isolated", "gpu_launches", "e2e")})
print("extras", json.dumps(d.get("extras"))[:600])
PY
