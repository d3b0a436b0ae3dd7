#!/bin/bash
# 2-GPU call (gpurun --gpus 2): bench.py under torchrun/NCCL for c2 (headline) and c3 (BASELINE configs[2], data-parallel
# spleen), and co-training iterations/s at N=2 (DDP) for c1/c3 -- launched exactly as the driver launches the scaling run.
tag=${1:-r03_n2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.csv 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
( timeout 300 $TR bench.py --gpus 2 2>&1 | tail -2 ) > $out/bench_c2_n2.log
( timeout 200 $TR bench.py --gpus 2 --workload c3 2>&1 | tail -2 ) > $out/bench_c3_n2.log
( timeout 120 $TR bench.py --gpus 2 --impl reference --steps 3 --warmup 1 2>&1 | tail -2 ) > $out/bench_ref_n2.log
for cfg in c3 c1; do
  ( timeout 300 $TR tools/cotrain_bench.py --config $cfg --arms ours,nets --out $out 2>&1 | tail -4 ) > $out/cotrain_${cfg}_n2.log
done
( timeout 200 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3_n1.log
cat $out/bench_c2_n2.log $out/bench_c3_n2.log $out/bench_ref_n2.log $out/cotrain_c3_n2.log $out/cotrain_c1_n2.log $out/bench_c3_n1.log
