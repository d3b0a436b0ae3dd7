#!/bin/bash
# 2-GPU call (gpurun --gpus 2): the fused peer exchange (loopback tests on GPU 0, then bench.py under torchrun exactly
# as the driver launches the scaling run, p2p vs NCCL exchange), c3 (BASELINE configs[2]) and co-training iterations/s
# at N=2 (DDP).  Full logs are kept (a crash must leave its traceback).
tag=${1:-r03_n2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpu.csv 2>&1
nvidia-smi topo -m > $out/topo.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python -m pytest tests/test_gpu_exchange.py -x -q -m gpu > $out/pytest_exchange.log 2>&1
timeout 300 $TR bench.py --gpus 2 > $out/bench_c2_n2_p2p.log 2>&1
timeout 300 $TR bench.py --gpus 2 --exchange nccl --e2e-steps 5 > $out/bench_c2_n2_nccl.log 2>&1
timeout 200 $TR bench.py --gpus 2 --workload c3 > $out/bench_c3_n2.log 2>&1
timeout 200 $TR bench.py --gpus 2 --workload c4 --steps 300 > $out/bench_c4_n2.log 2>&1
timeout 120 python bench.py --no-cpu-baseline > $out/bench_c2_n1.log 2>&1
for cfg in c3 c1; do
  timeout 300 $TR tools/cotrain_bench.py --config $cfg --arms ours,nets --out $out > $out/cotrain_${cfg}_n2.log 2>&1
done
tail -4 $out/pytest_exchange.log
for f in bench_c2_n2_p2p bench_c2_n2_nccl bench_c3_n2 bench_c4_n2 bench_c2_n1; do echo "== $f"; tail -1 $out/$f.log | cut -c1-900; done
for cfg in c3 c1; do echo "== cotrain $cfg"; grep -v "^\*\*\*\|OMP_NUM" $out/cotrain_${cfg}_n2.log | tail -25 | cut -c1-400; done
