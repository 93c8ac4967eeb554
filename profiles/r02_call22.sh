#!/bin/bash
# control-warp fused compress kernel (VKJIT_SCAN_CTRL=1): correctness subset, then geometries
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_SCAN_CTRL=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28" > $O/r02_c22_pytest.log 2>&1; tail -5 $O/r02_c22_pytest.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-330; }
{
run VKJIT_SCAN_CTRL=0
run VKJIT_SCAN_CTRL=1
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=1024 VKJIT_CTRL_VPT=2 VKJIT_CTRL_SLOTS=5
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=1024 VKJIT_CTRL_VPT=3 VKJIT_CTRL_SLOTS=4
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=512 VKJIT_CTRL_VPT=2 VKJIT_CTRL_SLOTS=4 VKJIT_CTRL_CTAS=2
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=1024 VKJIT_CTRL_VPT=2 VKJIT_CTRL_SLOTS=6 VKJIT_CTRL_DEPTH=3
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=1024 VKJIT_CTRL_VPT=4 VKJIT_CTRL_SLOTS=3
} 2>&1 | tee $O/r02_c22_ctrl_ab.txt
