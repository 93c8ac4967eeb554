#!/bin/bash
# pageable-upload staging ring: chunk size / copy threads sweep (e2e_default_numpy leg of bench.py)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('e2e', round(d['e2e']['value'],1), 'default_numpy', round(d['e2e_default_numpy']['value'],1), 'h2d', round(d['e2e_default_numpy']['h2d_GBps_per_gpu'],1))"; }
{
nproc
run VKJIT_STAGING_CHUNK_KB=2048
run VKJIT_STAGING_CHUNK_KB=8192
run VKJIT_STAGING_CHUNK_KB=16384
run VKJIT_STAGING_CHUNK_KB=8192 VKJIT_COPY_THREADS=12
run VKJIT_STAGING_CHUNK_KB=8192 VKJIT_COPY_THREADS=4
run VKJIT_STAGING_CHUNK_KB=4096 VKJIT_COPY_THREADS=6
run VKJIT_NO_STAGING=1
} 2>&1 | tee $O/r02_c37_staging.txt
