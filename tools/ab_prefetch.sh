#!/bin/bash
# A/B: L2 prefetch of a launch's first tiles before its dependency wait (DCT_TILE_PREFETCH = tiles per CTA, DCT_L2_PREFETCH)
out=${1:-gpurun_out/ab}; mkdir -p $out
for wl in c2 c3 c4 c1; do for cfg in "0 0" "2 0" "4 0" "8 0" "8 1" "4 1"; do set -- $cfg
  DCT_TILE_PREFETCH=$1 DCT_L2_PREFETCH=$2 timeout 300 python bench.py --workload $wl --steps 1200 --no-cpu-baseline --no-extras --e2e-steps 3 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl tile_prefetch=$1 l2_prefetch=$2 ms_per_step=%.4f stepGBps=%.0f | ' % (d['ms_per_step'], r['step_achieved_GBps']) + ' '.join('%s=%.2fus(%.0f%%)' % (k['part'], k['us'], 100*k['frac']) for k in r['step_kernels']))"
done; done | tee $out/ab_prefetch.log
for cfg in "0 0" "8 1"; do set -- $cfg; echo "== tile_prefetch=$1 l2_prefetch=$2"; DCT_TILE_PREFETCH=$1 DCT_L2_PREFETCH=$2 python tools/step_trace.py | tail -7; done | tee $out/trace_prefetch.log
