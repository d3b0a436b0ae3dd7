#!/bin/bash
# 1-GPU call r13: push-based cluster reduction in the perturbation normalisation: parity + c2 / c1 step time + launch list.
tag=${1:-r13}
out=gpurun_out/$tag
mkdir -p $out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $out/pytest_gpu.log
for rep in 1 2; do
  timeout 200 python bench.py --steps 3000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2 rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f' % (d['ms_per_step'], r['kernel_ms']*1e3))"
done > $out/c2_step.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/ncu_launches_bench_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c2.log 2>&1
tail -3 $out/pytest_gpu.log; cat $out/c2_step.log; grep "l2_cluster" $out/ncu_launches_bench_c2.csv | awk -F'","' '{print $NF}' | tr -d '"' | sort -n | head -3
