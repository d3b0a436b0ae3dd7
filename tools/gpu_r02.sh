#!/bin/bash
# Second GPU call of round 1: full GPU parity suite (incl. supervised / helper / ensemble kernels), smoke, bench (both
# arms), c4 bench, co-training iterations/s (1 GPU), ncu launch list of the bench and one full capture of the c4 JSD kernel.
tag=${1:-r02}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
nproc > $out/nproc.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
( timeout 600 python bench.py 2>&1 | tail -3 ) > $out/bench.log
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -3 ) > $out/bench_ref.log
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.log
for cfg in c1 c3 c2; do
  ( timeout 400 python tools/cotrain_bench.py --config $cfg --out $out 2>&1 | tail -8 ) > $out/cotrain_$cfg.log
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tile_kernel -s 4 -c 1 -o $out/prof_jsd_c4 \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_c4.log 2>&1
tail -5 $out/pytest_gpu.log; cat $out/smoke.log; cat $out/cotrain_c1.log $out/cotrain_c3.log $out/cotrain_c2.log; cat $out/bench.log $out/bench_ref.log $out/bench_c4.log
