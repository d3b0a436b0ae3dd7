#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench (both arms), ncu launch list + full capture of the top kernel.
# usage: gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
tag=${1:-r01}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
nproc > $out/nproc.txt
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
( timeout 600 python bench.py 2>&1 | tail -3 ) > $out/bench.log
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -3 ) > $out/bench_ref.log
( timeout 300 python bench.py --graph 0 --steps 500 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_nograph.log
for wl in c1 c3 c4; do
  ( timeout 300 python bench.py --workload $wl --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_$wl.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 6 -c 3 -o $out/prof_jsd \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 --graph 0 > $out/ncu_full.log 2>&1
( timeout 600 python tools/sweep.py --out $out/sweep 2>&1 | tail -12 ) > $out/sweep.log
ls -la $out
cat $out/pytest_gpu.log | tail -5; cat $out/smoke.log; cat $out/bench.log
