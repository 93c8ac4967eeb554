#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
echo "== default ctrl geometry (index mode: 2 x (512+32), L=2, D=3, window 320)"
VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -15
echo "== L=1 D=2 window 160"
VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=1 VKJIT_CTRL_DEPTH=2 VKJIT_LOOK_WIDE=5 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh_idx 2>&1 | tail -15
echo "== values, default ctrl geometry (1 x (512+32), S=5)"
VKJIT_SCAN_CTRL=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_ctrl_timeline.py thresh 2>&1 | tail -15
} > $O/r02_c24_ctrl_timeline.txt 2>&1
cat $O/r02_c24_ctrl_timeline.txt
