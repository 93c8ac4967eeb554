#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
timeout 400 python profiles/h2d_numa_probe.py 8 > gpurun_out/r02_c38_h2d_numa.txt 2>&1
cat gpurun_out/r02_c38_h2d_numa.txt | cut -c1-220
