#!/bin/bash
# status window through strong loads into registers (VKJIT_SCAN_WREG=1) vs cp.async.cg into shared memory
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
for w in 0 1; do
  echo "== VKJIT_SCAN_WREG=$w"
  VKJIT_SCAN_WREG=$w VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py thresh 2>&1 | tail -32
done
} > $O/r02_c20_fscan_timeline.txt 2>&1
cat $O/r02_c20_fscan_timeline.txt
{
echo "== WREG=0"; VKJIT_SCAN_WREG=0 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== WREG=1"; VKJIT_SCAN_WREG=1 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== WREG=1, T=1024 for the compress modes"; VKJIT_SCAN_WREG=1 VKJIT_SCAN_T=1024 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
} | tee $O/r02_c20_scan_ab.txt
VKJIT_SCAN_WREG=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 300 -k "scan or compress or prefix or C28 or fused or lagged" > $O/r02_c20_pytest.log 2>&1; tail -3 $O/r02_c20_pytest.log
