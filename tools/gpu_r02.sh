#!/bin/bash
# Second GPU call of round 1: parity of the new helper / ensemble kernels, co-training iterations/s (1 GPU), bench c2 + c4.
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
for cfg in c1 c3 c2; do
  ( timeout 600 python tools/cotrain_bench.py --config $cfg --out $out 2>&1 | tail -8 ) > $out/cotrain_$cfg.log
done
( timeout 600 python bench.py 2>&1 | tail -3 ) > $out/bench.log
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.log
tail -5 $out/pytest_gpu.log; cat $out/cotrain_c1.log $out/cotrain_c3.log $out/cotrain_c2.log; cat $out/bench.log $out/bench_c4.log
