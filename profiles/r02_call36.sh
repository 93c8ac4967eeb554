#!/bin/bash
# H26: bin-range passes (VKJIT_SADD_PASSES=1) against the shipped per-CTA privatisation
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
timeout 120 python profiles/hist_cluster_ab.py 2>&1 | tail -1
VKJIT_SADD_PASSES=1 timeout 120 python profiles/hist_cluster_ab.py 2>&1 | tail -1
VKJIT_SADD_PASSES=1 VKJIT_PASS_KB=64 timeout 120 python profiles/hist_cluster_ab.py 2>&1 | tail -1
VKJIT_SADD_PASSES=1 VKJIT_PASS_KB=96 timeout 120 python profiles/hist_cluster_ab.py 2>&1 | tail -1
} | tee $O/r02_c36_h26_passes.txt
VKJIT_SADD_PASSES=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 300 -k "scatter or H26 or hist" > $O/r02_c36_pytest.log 2>&1; tail -3 $O/r02_c36_pytest.log
