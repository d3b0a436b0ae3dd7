#!/bin/bash
# 1-GPU call r09 (final build of the round): K=4/K=3 C=19 shape check, GPU parity suite, smoke, bench (both arms, c1..c4),
# c5 sweep at 64Mi + 1Gi pixels, ncu launch list of the headline bench and a full capture of its JSD kernel.
tag=${1:-r09}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
nproc > $out/nproc.txt
( timeout 200 tools/kbench_reg 10 -1 16 1 7 2>&1 | grep -v trace ) > $out/kbench_wide_more.log
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
( timeout 600 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 ) > $out/bench_c2_reference_arm.json
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.json
( timeout 300 python bench.py --workload c1 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c1.json
( timeout 300 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
( timeout 400 python tools/sweep.py --no-aten --reps 5 --out $out/sweep_64Mi 2>&1 | tail -12 ) > $out/sweep_64Mi.log
( timeout 400 python tools/sweep.py --no-aten --reps 2 --pixels 1073741824 --ks 2,3,4 --cs 4,19 --mem-gb 60 --out $out/sweep_1Gi 2>&1 | tail -8 ) > $out/sweep_1Gi.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/ncu_launches_bench_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:JsdOp -s 4 -c 1 -o $out/prof_jsd_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_c2.log 2>&1
ncu -i $out/prof_jsd_c2.ncu-rep --page raw --csv > $out/ncu_full_raw_c2_jsd.csv 2>/dev/null
ncu -i $out/prof_jsd_c2.ncu-rep --page details > $out/ncu_full_details_c2_jsd.txt 2>/dev/null
rm -f $out/prof_jsd_c2.ncu-rep
cat $out/kbench_wide_more.log; tail -3 $out/pytest_gpu.log; cat $out/smoke.log; cut -c1-330 $out/bench_c2_default.json; cut -c1-200 $out/bench_c4.json; cat $out/sweep_64Mi.log | cut -c1-200
