#!/bin/bash
# 1-GPU call r10: parity suite with the final K=4 launch shape, C=19 sweep rows, the c2-sized copy ceiling of the tile
# pipeline itself (tools/kbench_tile which=0 with its in-kernel trace).
tag=${1:-r10}
out=gpurun_out/$tag
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 ) > $out/pytest_gpu.log
( timeout 400 python tools/sweep.py --no-aten --reps 5 --cs 19 --out $out/sweep_64Mi_c19 2>&1 | tail -4 ) > $out/sweep_64Mi_c19.log
( timeout 300 tools/kbench_reg 20 -1 32 1 0 2>&1 ) > $out/kbench_c2_family.log
tail -3 $out/pytest_gpu.log; cut -c1-200 $out/sweep_64Mi_c19.log; grep -A1 "jsd+dice c2 \|copy3x4\|klfromlogits\|kllogit" $out/kbench_c2_family.log | cut -c1-220 | head -80
