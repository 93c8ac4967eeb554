#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { echo "== T=$1 VPT=$2 SLOTS=$3 LAGMAX=$4"; VKJIT_SCAN_T=$1 VKJIT_SCAN_VPT=$2 VKJIT_SCAN_SLOTS=$3 VKJIT_LAG_MAX_NODES=$4 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1; }
{
echo "== default"; timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
run 512 4 3 8
run 512 2 4 8
run 512 4 3 64
run 1024 2 4 8
} | tee $O/r02_c13_fused_scan_T.txt
VKJIT_SCAN_T=512 VKJIT_SCAN_VPT=4 VKJIT_SCAN_SLOTS=3 timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout 200 -k "lagged or fused" > $O/r02_c13_pytest.log 2>&1; tail -3 $O/r02_c13_pytest.log
