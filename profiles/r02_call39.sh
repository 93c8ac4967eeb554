#!/bin/bash
# parked-result lagged prefix sum for traces without streamed inputs (VKJIT_SCAN_PARK=1)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | grep -o "scan_trace [^)]*)"; }
{
run VKJIT_SCAN_PARK=0
run VKJIT_SCAN_PARK=1
run VKJIT_SCAN_PARK=1 VKJIT_PARK_T=512
run VKJIT_SCAN_PARK=1 VKJIT_PARK_VPT=4
run VKJIT_SCAN_PARK=1 VKJIT_PARK_T=512 VKJIT_PARK_VPT=8
run VKJIT_SCAN_PARK=1 VKJIT_PARK_VPT=7
} 2>&1 | tee $O/r02_c39_park.txt
VKJIT_SCAN_PARK=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 300 -k "prefix or scan or fused or lagged or C28" > $O/r02_c39_pytest.log 2>&1; tail -3 $O/r02_c39_pytest.log
