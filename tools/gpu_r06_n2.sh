#!/bin/bash
# 2-GPU call r06: final-build bench under torchrun exactly as the driver's scaling run launches it (default exchange =
# chained p2p), A/B against NCCL and the fused variant, c3 / c4, N=1 on the same box, co-training iterations/s at N=2.
tag=${1:-r06_n2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.csv 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 200 python bench.py --no-cpu-baseline > $out/bench_c2_n1.log 2>&1
timeout 300 $TR bench.py --gpus 2 > $out/bench_c2_n2_p2p.log 2>&1
timeout 300 $TR bench.py --gpus 2 --exchange nccl --e2e-steps 5 > $out/bench_c2_n2_nccl.log 2>&1
timeout 300 $TR bench.py --gpus 2 --exchange p2p-fused --e2e-steps 5 > $out/bench_c2_n2_p2p_fused.log 2>&1
timeout 200 python bench.py --workload c3 --no-cpu-baseline > $out/bench_c3_n1.log 2>&1
timeout 200 $TR bench.py --gpus 2 --workload c3 > $out/bench_c3_n2.log 2>&1
timeout 200 python bench.py --workload c4 --steps 300 --no-cpu-baseline > $out/bench_c4_n1.log 2>&1
timeout 200 $TR bench.py --gpus 2 --workload c4 --steps 300 > $out/bench_c4_n2.log 2>&1
for cfg in c3 c1; do
  timeout 300 $TR tools/cotrain_bench.py --config $cfg --arms ours,nets --out $out > $out/cotrain_${cfg}_n2.log 2>&1
done
for f in bench_c2_n1 bench_c2_n2_p2p bench_c2_n2_nccl bench_c2_n2_p2p_fused bench_c3_n1 bench_c3_n2 bench_c4_n1 bench_c4_n2; do tail -1 $out/$f.log | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$f', d['n_gpus'], 'ms/step %.4f' % d['ms_per_step'], 'Gpix/s %.2f' % (d['value']/1e9), 'e2e %.1f' % (d['e2e']['value']/1e6), d['config'].get('exchange','')[:30], d['config'].get('exchange_check'))
except Exception as e: print('$f ERR', e)
"; done
for cfg in c3 c1; do grep '^{"metric"' $out/cotrain_${cfg}_n2.log | cut -c1-200; done
