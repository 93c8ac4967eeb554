#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 900 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py tests/test_frontend_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r02_c50_pytest.log 2>&1; tail -3 gpurun_out/r02_c50_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
