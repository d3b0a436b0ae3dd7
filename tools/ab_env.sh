#!/bin/bash
# A/B of the tile schedule's environment knobs at the product launch shapes: DCT_TILE_POOL_DIV (tail pool = 1/div of a CTA's range),
# DCT_TILE_PREFETCH (tiles per CTA prefetched into L2 before the dependency wait).  Every run is bounded (60 s).
out=${1:-gpurun_out/ab}; mkdir -p $out
line() { python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 ms_per_step=%.4f stepGBps=%.0f | ' % (d['ms_per_step'], r['step_achieved_GBps']) + ' '.join('%s=%.2fus(%.0f%%)' % (k['part'], k['us'], 100*k['frac']) for k in r['step_kernels']))"; }
for wl in c2 c3; do
  for pd in 0 2 3 4 8 16; do DCT_TILE_POOL_DIV=$pd timeout 60 python bench.py --workload $wl --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | line "$wl pool_div=$pd"; done
  for pf in 0 1 3 4; do DCT_TILE_PREFETCH=$pf timeout 60 python bench.py --workload $wl --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | line "$wl prefetch=$pf"; done
done | tee $out/ab_env.log
