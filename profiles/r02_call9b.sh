#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -x -q --timeout 500 > gpurun_out/r02_c9b_pytest_mgpu.log 2>&1; tail -6 gpurun_out/r02_c9b_pytest_mgpu.log
