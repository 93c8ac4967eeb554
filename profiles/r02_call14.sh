#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
echo "== default"; timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== non-lag T=512"; VKJIT_SCAN_T_NOLAG=512 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
} | tee $O/r02_c14_fused_scan.txt
VKJIT_SCAN_T_NOLAG=512 timeout 300 python -m pytest tests/test_cuda_parity.py -m gpu -x -q --timeout 200 -k "lagged or fused" > $O/r02_c14_pytest.log 2>&1; tail -3 $O/r02_c14_pytest.log
timeout 300 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 200 -k "lagged or fused or C28 or compress" > $O/r02_c14_pytest2.log 2>&1; tail -3 $O/r02_c14_pytest2.log
