#!/bin/bash
# ncu counters of the Monte-Carlo kernel with and without the range-proven fast paths
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
M=smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,gpu__time_duration.sum,smsp__inst_executed_per_warp.ratio,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_xu.sum,smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct,smsp__warp_issue_stalled_no_instruction_per_warp_active.pct,smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct,smsp__warp_issue_stalled_not_selected_per_warp_active.pct,smsp__warp_issue_stalled_wait_per_warp_active.pct
timeout 300 ncu --clock-control none -k regex:vkjit_trace -s 1 -c 1 --metrics $M --csv --log-file $O/r02_c35_m26_ranges.csv python profiles/prof_m26.py > $O/r02_c35_a.log 2>&1
VKJIT_NO_RANGES=1 timeout 300 ncu --clock-control none -k regex:vkjit_trace -s 1 -c 1 --metrics $M --csv --log-file $O/r02_c35_m26_noranges.csv python profiles/prof_m26.py > $O/r02_c35_b.log 2>&1
grep -h "vkjit_trace" $O/r02_c35_m26_ranges.csv | cut -d, -f13-15 | tr -d '"'
echo ---; grep -h "vkjit_trace" $O/r02_c35_m26_noranges.csv | cut -d, -f13-15 | tr -d '"'
