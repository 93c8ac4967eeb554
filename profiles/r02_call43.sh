#!/bin/bash
# control-warp kernel v6: packed aggregate window (5 coalesced cp.async per lane instead of 320 status lines)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_SCAN_CTRL=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28" > $O/r02_c43_pytest.log 2>&1; tail -3 $O/r02_c43_pytest.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_SCAN_CTRL=1
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=5
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=5
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7
run VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=2 VKJIT_CTRL_DEPTH=3 VKJIT_FSCAN_DIAG=4
} 2>&1 | tee $O/r02_c43_ctrl_ab.txt
{
echo "== index mode L=3 D=4"; VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=4 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -14
} > $O/r02_c43_ctrl_timeline.txt 2>&1
cat $O/r02_c43_ctrl_timeline.txt
