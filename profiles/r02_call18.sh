#!/bin/bash
# early status-window prefetch in the lagged kernels (fused: VKJIT_SCAN_EARLY, default on; hand-written: VKJIT_SCAN_EARLY_PRIM=1)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
{
for e in 0 1; do for c in thresh hash_mask; do
  echo "== VKJIT_SCAN_EARLY=$e"
  VKJIT_SCAN_EARLY=$e VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py $c 2>&1 | tail -12
done; done
} > $O/r02_c18_fscan_timeline.txt 2>&1
cat $O/r02_c18_fscan_timeline.txt
{
echo "== fused, early (default)"; timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== fused, VKJIT_SCAN_EARLY=0"; VKJIT_SCAN_EARLY=0 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== fused, early, T=1024 for the compress modes"; VKJIT_SCAN_T=1024 timeout 150 python profiles/fused_scan_ab.py 2>&1 | tail -1
echo "== hand-written, default"; timeout 150 python profiles/scan_ab.py 2>&1 | tail -1
echo "== hand-written, VKJIT_SCAN_EARLY_PRIM=1"; VKJIT_SCAN_EARLY_PRIM=1 timeout 150 python profiles/scan_ab.py 2>&1 | tail -1
} | tee $O/r02_c18_scan_ab.txt
timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 300 -k "scan or compress or prefix or C28 or fused or lagged" > $O/r02_c18_pytest.log 2>&1; tail -3 $O/r02_c18_pytest.log
VKJIT_SCAN_EARLY_PRIM=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 300 -k "scan or compress or prefix or C28" > $O/r02_c18_pytest_prim.log 2>&1; tail -3 $O/r02_c18_pytest_prim.log
