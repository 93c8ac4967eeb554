#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py thresh > gpurun_out/r02_c44_fscan_timeline.txt 2>&1
tail -24 gpurun_out/r02_c44_fscan_timeline.txt
