#!/bin/bash
# control-warp fused compress kernel v5: warp rows leave through a staging row as aligned coalesced stores
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_SCAN_CTRL=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28" > $O/r02_c29_pytest.log 2>&1; tail -5 $O/r02_c29_pytest.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7 VKJIT_FSCAN_DIAG=2
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=5
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=2 VKJIT_CTRL_DEPTH=3
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=5 VKJIT_CTRL_VPT=2
} 2>&1 | tee $O/r02_c29_ctrl_ab.txt
