#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
echo "== index mode, L=6 D=7, lane-by-lane stores"; VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7 VKJIT_FSCAN_DIAG=2 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -14
echo "== index mode, L=6 D=7, no stores"; VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7 VKJIT_FSCAN_DIAG=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -14
} > $O/r02_c30_ctrl_timeline.txt 2>&1
cat $O/r02_c30_ctrl_timeline.txt
