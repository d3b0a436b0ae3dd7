#!/bin/bash
# 1-GPU call r14: A/B of the normalisation kernel's exit (trailing cluster barrier / early PDL trigger) on the c2 step
out=gpurun_out/${1:-r14}; mkdir -p $out
for rep in 1 2; do for v in 0 1 2 3; do
  DCT_B200_L2_VARIANT=$v timeout 200 python bench.py --steps 3000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('variant=$v rep$rep ms_per_step=%.4f' % d['ms_per_step'])"
done; done | tee $out/ab_l2_exit.log
