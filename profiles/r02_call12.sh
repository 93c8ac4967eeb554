#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout 200 -k "hot_bins or privatised" > $O/r02_c12_pytest.log 2>&1; tail -5 $O/r02_c12_pytest.log
for v in 0 1; do VKJIT_SADD_CLUSTER=$v timeout 120 python profiles/hist_cluster_ab.py 2>&1 | tail -1; done | tee $O/r02_c12_hist_cluster.txt
