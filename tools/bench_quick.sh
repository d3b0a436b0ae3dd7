#!/bin/bash
# per-launch times of every workload's step (no extras, no CPU baseline); DCT_* knobs pass through the environment
out=${1:-gpurun_out/q}; mkdir -p $out
for wl in c2 c3 c1 c4; do for rep in 1 2; do
  timeout 300 python bench.py --workload $wl --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl ms_per_step=%.4f stepGBps=%.0f (%.1f%%) | ' % (d['ms_per_step'], r['step_achieved_GBps'], 100*r['step_achieved_GBps']/r['peak']) + ' '.join('%s=%.2fus(%.0f%%)' % (k['part'], k['us'], 100*k['frac']) for k in r['step_kernels']))"
done; done | tee $out/bench_quick.log
