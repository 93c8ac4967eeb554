#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1 | cut -c1-215; }
{
run VKJIT_FSCAN_DIAG=1
run VKJIT_FSCAN_DIAG=1 VKJIT_SCAN_CTRL=1 VKJIT_CTRL_LAG=6 VKJIT_CTRL_DEPTH=7
} 2>&1 | tee $O/r02_c28_diag.txt
