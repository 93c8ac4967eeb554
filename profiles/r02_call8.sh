#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
timeout 400 python -m pytest tests/test_cuda_parity.py tests/test_cuda_fullsize.py -m gpu -x -q --timeout 120 > $O/r02_c8_pytest.log 2>&1; tail -5 $O/r02_c8_pytest.log
timeout 120 python profiles/fused_scan_ab.py > $O/r02_c8_fused_default.txt 2>&1; tail -1 $O/r02_c8_fused_default.txt
VKJIT_LAG_MAX_NODES=64 timeout 120 python profiles/fused_scan_ab.py > $O/r02_c8_fused_lag64.txt 2>&1; tail -1 $O/r02_c8_fused_lag64.txt
python - <<'PY' > gpurun_out/r02_c8_hot.txt 2>&1
import os, subprocess, sys, json
code = r"""
import sys, json, numpy as np
sys.path.insert(0, '.')
import torch, vkjit_b200 as vk
from bench import hash_trace
from vkjit_b200.ir import Bop, Ir, VarType as T
vk.init(0)
stream = torch.cuda.ExternalStream(vk.stream_ptr())
ir = Ir(); c = ir.const_u32
m = 1 << 26
lanes = ir.arange(T.U32, m)
res = {}
for nb in (1, 16, 256, 4096, 65536):
    idx = ir.bop(Bop.And, hash_trace(ir, lanes, 77), c(nb - 1)); ir.eval([idx])
    bins = ir.array_u32(np.zeros(nb, np.uint32)); one = c(1)
    ts = []
    for i in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); s = ir.scatter_add(one, bins, idx); ir.eval([s]); b.record(stream); vk.sync(); ir.dec_ref_count(s)
        if i >= 2: ts.append(a.elapsed_time(b))
    assert int(ir.as_slice(bins, T.U32).astype(np.uint64).sum()) == 6 * m
    res[nb] = round(sorted(ts)[len(ts)//2], 4)
    ir.dec_ref_count(idx); ir.dec_ref_count(bins)
print(json.dumps(res))
"""
for env in ({}, {"VKJIT_NO_AGG": "1"}):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **env), capture_output=True, text=True, timeout=150)
    print("count histogram ms by number of bins, 2^26 lanes", env or "default (probe)", r.stdout.strip(), r.stderr[-300:])
PY
cat gpurun_out/r02_c8_hot.txt
