#!/bin/bash
# 1-GPU call r21 (last of the round, ~3 GPU-minutes left): ncu launch list of the headline bench on the final build.
out=gpurun_out/${1:-r21}; mkdir -p $out
timeout 110 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_bench_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c2.log 2>&1
grep -c tile_kernel $out/ncu_launches_bench_c2.csv
( timeout 60 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
cut -c1-200 $out/bench_c3.json
