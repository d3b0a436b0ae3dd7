#!/bin/bash
# A/B of the step time: PDL on/off x CUDA graph on/off (prints ms_per_step and the JSD kernel time)
out=gpurun_out/${1:-ab}; mkdir -p $out
for pdl in 1 0; do for g in 1 0; do
  DCT_B200_PDL=$pdl timeout 200 python bench.py --steps 1000 --graph $g --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('pdl=$pdl graph=$g ms_per_step=%.4f kernel_us=%.2f frac=%.3f stepGBps=%.0f' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac'], r['step_achieved_GBps']))"
done; done | tee $out/ab.log
