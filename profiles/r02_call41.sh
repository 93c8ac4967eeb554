#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 900 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout 300 -k "bin_range_passes or hot_bins or kernel_variants or ranged_random or staging_ring" > $O/r02_c41_pytest.log 2>&1; tail -5 $O/r02_c41_pytest.log
