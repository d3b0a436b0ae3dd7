#!/bin/bash
# A/B of the perturbation-normalisation launch: cluster kernels (DCT_L2_VARIANT=1) vs the co-resident one-launch kernel
# (default), per-kernel times from bench.py's step_kernels.
out=${1:-gpurun_out/ab}; mkdir -p $out
for wl in c2 c3 c1; do for v in 1 0; do
  DCT_L2_VARIANT=$v timeout 300 python bench.py --workload $wl --steps 1000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl l2_variant=$v ms_per_step=%.4f stepGBps=%.0f | ' % (d['ms_per_step'], r['step_achieved_GBps']) + ' '.join('%s=%.2fus(%.0f%%)' % (k['part'], k['us'], 100*k['frac']) for k in r['step_kernels']))"
done; done | tee $out/ab_l2.log
