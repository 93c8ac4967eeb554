#!/bin/bash
# control-warp fused compress kernel v2 (control-level lag L, window requested one iteration ahead)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_SCAN_CTRL=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28" > $O/r02_c23_pytest.log 2>&1; tail -5 $O/r02_c23_pytest.log
VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=2 timeout 600 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged" > $O/r02_c23_pytest2.log 2>&1; tail -3 $O/r02_c23_pytest2.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_SCAN_CTRL=0
run VKJIT_SCAN_CTRL=1
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=2
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=2 VKJIT_LOOK_WIDE=5
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=2 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4 VKJIT_CTRL_SLOTS=6
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=4 VKJIT_CTRL_LAG=1 VKJIT_CTRL_DEPTH=2 VKJIT_CTRL_SLOTS=4
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_T=256 VKJIT_CTRL_VPT=4 VKJIT_CTRL_CTAS=2
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_VPT=4 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4 VKJIT_CTRL_SLOTS=6
} 2>&1 | tee $O/r02_c23_ctrl_ab.txt
