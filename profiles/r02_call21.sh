#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py thresh > $O/r02_c21_fscan_timeline.txt 2>&1
tail -12 $O/r02_c21_fscan_timeline.txt
