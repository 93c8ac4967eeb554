#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for c in thresh hash_mask thresh_idx; do
  VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py $c 2>&1 | tail -14
done | tee $O/r02_c17_fscan_timeline.txt
