#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for SP in 0 1 2 4; do
for L in 25 28; do
VKJIT_REDUCE_SPARE_CTAS=$SP VKJIT_REDUCE_TRACE=1 python profiles/reduce_timeline.py --steps 24 --log2n $L --arrays 4 --out $O/r02_sp${SP}_$L > $O/r02_sp${SP}_$L.json 2> $O/r02_sp${SP}_$L.err
done
done
python bench.py --steps 20 --warmup 5 --no-extras > $O/r02_c5_bench_n1.json 2> $O/r02_c5_bench_n1.err
VKJIT_REDUCE_SPARE_CTAS=0 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > $O/r02_c5_bench_n1_sp0.json 2> $O/r02_c5_bench_n1_sp0.err
tail -c 300 $O/r02_sp2_25.err
