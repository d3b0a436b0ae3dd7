#!/bin/bash
# C=19 investigation: kbench sweeps (copy ceiling vs JSD/KL ops) and one ncu --set full capture of the c4 JSD kernel
out=gpurun_out/${1:-c19}; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.csv
( timeout 300 tools/kbench_tile 20 -1 16 1 1 2>&1 ) > $out/kbench_which1.log
( timeout 300 tools/kbench_tile 20 -1 16 1 2 2>&1 ) > $out/kbench_which2.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -o $out/prof_j19 \
    tools/kbench_tile 3 3 16 1 1 > $out/ncu_j19.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 1 -o $out/prof_copy19 \
    tools/kbench_tile 3 8 16 1 1 > $out/ncu_copy19.log 2>&1
tail -30 $out/kbench_which1.log
