#!/bin/bash
# A/B: refill of a drained stage right after its own store group has been read (DCT_TILE_REFILL=1) vs one tile later (0); unset = product rule
out=${1:-gpurun_out/ab}; mkdir -p $out
for wl in c2 c3 c4; do for r in "" 0 1 "" 0 1; do
  DCT_TILE_REFILL=$r timeout 60 python bench.py --workload $wl --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl refill=${r:-rule} ms_per_step=%.4f stepGBps=%.0f | ' % (d['ms_per_step'], r['step_achieved_GBps']) + ' '.join('%s=%.2fus(%.0f%%)' % (k['part'], k['us'], 100*k['frac']) for k in r['step_kernels']))"
done; done | tee $out/ab_refill.log
