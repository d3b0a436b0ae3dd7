#!/bin/bash
# 1-GPU call r07: GPU parity suite (adds the shared-memory-resident wide JSD body, deferred publication), C=19 sweep rows,
# exchange variants in loopback incl. the deferred (forked-branch) publication.
tag=${1:-r07}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 400 python tools/sweep.py --no-aten --reps 5 --cs 19 --out $out/sweep_64Mi_c19 2>&1 | tail -4 ) > $out/sweep_64Mi_c19.log
for rep in 1 2; do
  for ex in auto p2p p2p-chained; do
    timeout 200 python bench.py --workload c2 --steps 3000 --no-cpu-baseline --e2e-steps 5 --exchange $ex 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2 exchange=$ex rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f launches/step=%d check=%s' % (d['ms_per_step'], r['kernel_ms']*1e3, d['gpu_launches']//d['steps'], d['config'].get('exchange_check')))"
  done
done > $out/ab_exchange_loopback.log 2>&1
tail -4 $out/pytest_gpu.log; cat $out/sweep_64Mi_c19.log | cut -c1-230; cat $out/ab_exchange_loopback.log
