#!/bin/bash
# ncu --set full of the shipped lagged fused compress kernel (compress_values(v, v > t), 2^28 lanes)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:vkjit_trace -s 1 -c 1 -o $O/r02_c40_fused_compress python profiles/prof_fused_scan.py > $O/r02_c40.log 2>&1; tail -2 $O/r02_c40.log
ncu -i $O/r02_c40_fused_compress.ncu-rep --page raw --csv > $O/r02_c40_raw.csv 2>/dev/null; wc -c $O/r02_c40_raw.csv
