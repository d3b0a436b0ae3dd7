#!/bin/bash
# 1-GPU call r08: K*C > 40 launch-shape sweep, register-resident vs shared-memory-resident JSD body (tools/kbench_tile which=6),
# and the publication kernel (one thread per value x peer) in loopback.
tag=${1:-r08}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.csv 2>&1
( timeout 300 tools/kbench_stream 10 -1 16 1 6 2>&1 | grep -v trace ) > $out/kbench_wide_stream.log
( timeout 300 tools/kbench_reg 10 -1 16 1 6 2>&1 | grep -v trace ) > $out/kbench_wide_reg.log
for rep in 1 2; do
  for ex in auto p2p; do
    timeout 200 python bench.py --workload c2 --steps 3000 --no-cpu-baseline --e2e-steps 5 --exchange $ex 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('c2 exchange=$ex rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f launches/step=%d check=%s' % (d['ms_per_step'], r['kernel_ms']*1e3, d['gpu_launches']//d['steps'], d['config'].get('exchange_check')))"
  done
done > $out/ab_exchange_loopback.log 2>&1
( timeout 300 python -m pytest tests/test_gpu_exchange.py -m gpu -x -q 2>&1 | tail -3 ) > $out/pytest_exchange.log
echo "== stream"; cat $out/kbench_wide_stream.log; echo "== reg"; cat $out/kbench_wide_reg.log; cat $out/ab_exchange_loopback.log; cat $out/pytest_exchange.log
