#!/bin/bash
# control-warp fused compress kernel v4: single-reduction anchored walk, TMA producer warp
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_SCAN_CTRL=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28" > $O/r02_c26_pytest.log 2>&1; tail -5 $O/r02_c26_pytest.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_SCAN_CTRL=1
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4 VKJIT_CTRL_SLOTS=6
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=5 VKJIT_CTRL_VPT=2
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=1 VKJIT_CTRL_DEPTH=2
} 2>&1 | tee $O/r02_c26_ctrl_ab.txt
{
echo "== index mode"; VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -14
echo "== values"; VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh 2>&1 | tail -14
} > $O/r02_c26_ctrl_timeline.txt 2>&1
cat $O/r02_c26_ctrl_timeline.txt
