#!/bin/bash
# 1-GPU call r17: cluster of 16 CTAs for 512 KB..1 MB samples (c3: 512 x 512 slices): parity, c3 step, launch list.
out=gpurun_out/${1:-r17}; mkdir -p $out
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $out/pytest_gpu.log
for rep in 1 2; do
  timeout 200 python bench.py --workload c3 --steps 3000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('c3 rep$rep ms_per_step=%.4f launches/step=%d' % (d['ms_per_step'], d['gpu_launches']//d['steps']))"
done > $out/c3_step.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/ncu_launches_bench_c3.csv \
    python bench.py --workload c3 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c3.log 2>&1
( timeout 200 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
tail -3 $out/pytest_gpu.log; cat $out/c3_step.log; grep "l2_" $out/ncu_launches_bench_c3.csv | awk -F'","' '{print $5, $NF}' | cut -c1-60,200- | sort | uniq -c | sort -rn | head -8
