#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
O=gpurun_out
for L in 25 26 28; do
VKJIT_REDUCE_TRACE=1 python profiles/reduce_timeline.py --steps 24 --log2n $L --arrays 8 --out $O/r02_tl1_$L > $O/r02_tl1_$L.json 2> $O/r02_tl1_$L.err
done
VKJIT_REDUCE_TRACE=1 python profiles/reduce_timeline.py --steps 24 --log2n 25 --arrays 1 --out $O/r02_tl1_25_a1 > $O/r02_tl1_25_a1.json 2> $O/r02_tl1_25_a1.err
VKJIT_REDUCE_TRACE=1 VKJIT_REDUCE_OVERLAP=0 python profiles/reduce_timeline.py --steps 24 --log2n 25 --arrays 8 --out $O/r02_tl1_25_nopdl > $O/r02_tl1_25_nopdl.json 2> $O/r02_tl1_25_nopdl.err
tail -c 300 $O/r02_tl1_25.err
