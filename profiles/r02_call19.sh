#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
for lw in 5; do for c in thresh; do
  echo "== VKJIT_LOOK_WIDE=$lw"
  VKJIT_LOOK_WIDE=$lw VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py $c 2>&1 | tail -32
done; done
} > $O/r02_c19_fscan_timeline.txt 2>&1
cat $O/r02_c19_fscan_timeline.txt
