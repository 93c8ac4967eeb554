#!/bin/bash
# Round 2, GPU call 1 (one B200): measurement record of the code as it entered the round + sanity.
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > $O/r02_c1_clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 5 > $O/r02_c1_bench_n1.json 2> $O/r02_c1_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02_c1_bench_ref.json 2> $O/r02_c1_bench_ref.err
kill $SMI
python -m pytest tests -m gpu -x -q > $O/r02_c1_pytest.log 2>&1
tail -3 $O/r02_c1_pytest.log
# launch list of the CURRENT bench loop (shares only; never a bench number)
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02_c1_launches.csv \
    python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > $O/r02_c1_ncu_bench.log 2>&1
# H26 gather + scatter-add: full capture (atomic counters live in the memory tables)
PROF_LOG2N=22 ncu --set full --clock-control none --import-source on -k regex:vkjit_trace -s 3 -c 1 -o $O/r02_c1_h26 \
    python profiles/prof_kernels.py hist > $O/r02_c1_ncu_h26.log 2>&1
head -c 1500 $O/r02_c1_bench_n1.json
