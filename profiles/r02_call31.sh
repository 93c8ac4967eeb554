#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=2 VKJIT_CTRL_LAG=2 VKJIT_CTRL_DEPTH=4
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=2 VKJIT_CTRL_LAG=3 VKJIT_CTRL_DEPTH=5
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=2 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=6
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=2 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=8
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=2 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=9
run VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_DIAG=0 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=6
} 2>&1 | tee $O/r02_c31_ctrl_ab.txt
{
echo "== index mode, L=4 D=6, lane-by-lane stores"; VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=4 VKJIT_CTRL_DEPTH=6 VKJIT_FSCAN_DIAG=2 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -14
} > $O/r02_c31_ctrl_timeline.txt 2>&1
cat $O/r02_c31_ctrl_timeline.txt
