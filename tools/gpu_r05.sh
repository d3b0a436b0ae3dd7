#!/bin/bash
# 1-GPU call r05: GPU parity suite (adds CoTrainStep vs torch, chained exchange), A/B on one box of the exchange
# variants in loopback {none, fused *_pub kernel, PDL-chained publication kernel} for c2 and c4, C=19 sweep rows
# re-measured with the clean build, co-training iterations/s for the Cityscapes flavour (IoU meter fused into the loss).
tag=${1:-r05}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
for wl in c2 c4; do
  steps=3000; [ $wl = c4 ] && steps=300
  for rep in 1 2; do
    for ex in auto p2p p2p-chained; do
      timeout 200 python bench.py --workload $wl --steps $steps --no-cpu-baseline --e2e-steps 5 --exchange $ex 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl exchange=$ex rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f frac=%.3f launches/step=%d' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac'], d['gpu_launches']//d['steps']))"
    done
  done
done > $out/ab_exchange_loopback.log 2>&1
( timeout 400 python tools/sweep.py --no-aten --reps 5 --cs 19 --out $out/sweep_64Mi_c19 2>&1 | tail -4 ) > $out/sweep_64Mi_c19.log
( timeout 400 python tools/sweep.py --no-aten --reps 2 --pixels 1073741824 --ks 2,3 --cs 19 --mem-gb 60 --out $out/sweep_1Gi_c19 2>&1 | tail -3 ) > $out/sweep_1Gi_c19.log
( timeout 400 python tools/cotrain_bench.py --config c4 --arms ours,nets --iters 30 --out $out 2>&1 | tail -4 ) > $out/cotrain_c4.log
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.json
tail -3 $out/pytest_gpu.log; cat $out/ab_exchange_loopback.log; cat $out/sweep_64Mi_c19.log $out/sweep_1Gi_c19.log | cut -c1-220; cut -c1-400 $out/cotrain_c4.log
