#!/bin/bash
# shipped lagged fused kernels with the packed, anchored look-back (VKJIT_LAG_PACKED=1)
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
VKJIT_LAG_PACKED=1 timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 -k "compress or fused or lagged or C28 or prefix or scan" > $O/r02_c49_pytest.log 2>&1; tail -3 $O/r02_c49_pytest.log
run() { echo "== $*"; env "$@" timeout 120 python profiles/fused_scan_ab.py 2>&1 | tail -1; }
{
run VKJIT_LAG_PACKED=0
run VKJIT_LAG_PACKED=1
run VKJIT_LAG_PACKED=1 VKJIT_SCAN_T=1024
} 2>&1 | tee $O/r02_c49_lagpack.txt
VKJIT_LAG_PACKED=1 VKJIT_FSCAN_TRACE=/tmp/fscan.bin timeout 200 python profiles/fscan_timeline.py thresh 2>&1 | sed -n 1,20p > $O/r02_c49_timeline.txt; cat $O/r02_c49_timeline.txt
