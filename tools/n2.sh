#!/bin/bash
# both bench arms under torch.distributed.run as the driver launches them (N = number of visible GPUs)
out=${1:-gpurun_out/n}; mkdir -p $out
N=$(nvidia-smi -L | wc -l)
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --impl reference --steps 3 --warmup 1 2>$out/bench_ref_n$N.err | tail -1 ) > $out/bench_ref_n$N.json; cut -c1-300 $out/bench_ref_n$N.json
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N 2>$out/bench_n$N.err | tail -1 ) > $out/bench_c2_n$N.json; cut -c1-1200 $out/bench_c2_n$N.json
( timeout 300 python -m pytest tests/test_gpu_exchange.py -m gpu -q 2>&1 | tail -3 )
