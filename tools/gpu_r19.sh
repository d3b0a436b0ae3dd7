#!/bin/bash
# 1-GPU call r19: Dice counters kept in the consumers' registers across tiles (DCT_DICE_LOCAL=1, the new default) against
# the per-tile fold through the producer warp (=0): parity suite on the new build, kernel A/B (tools/kbench_tile.cu built
# both ways), step A/B (DCT_B200_LIB = the old-scheme build), then the headline bench line + ncu launch list / full capture.
tag=${1:-r19}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
# kernel A/B: config indices of which=0 -- 0 jsd+dice c2 (product shape), 8 jsd c2 without Dice, 49 Dice meter alone
# (4 px/thread), 57 jsd+dice c3 (4 px/thread), 64 jsd+dice c1 x 8
for rep in 1 2; do for idx in 0 8 49 57 64; do for v in fold local; do
  echo -n "$v rep$rep "; timeout 60 tools/kbench_dice_$v 30 $idx 32 1 0 2>&1 | grep -v trace
done; done; done > $out/kbench_dice_ab.log 2>&1
cat $out/kbench_dice_ab.log | cut -c1-150
step() {  # $1 = label, $2 = workload, env DCT_B200_LIB
  timeout 200 python bench.py --workload $2 --steps 2000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 $2 ms_per_step=%.4f kernel_us=%.2f frac=%.3f' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac']))"
}
for rep in 1 2; do
  DCT_B200_LIB=$PWD/tools/ab/libdct_b200_fold.so step fold c2
  step local c2
done > $out/ab_step.log 2>&1
DCT_B200_LIB=$PWD/tools/ab/libdct_b200_fold.so step fold c1 >> $out/ab_step.log 2>&1
step local c1 >> $out/ab_step.log 2>&1
cat $out/ab_step.log
( timeout 400 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json
cut -c1-400 $out/bench_c2_default.json
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
cat $out/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/ncu_launches_bench_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:JsdOp -s 4 -c 1 -o $out/prof_jsd_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_c2.log 2>&1
ncu -i $out/prof_jsd_c2.ncu-rep --page raw --csv > $out/ncu_full_raw_c2_jsd.csv 2>/dev/null
ncu -i $out/prof_jsd_c2.ncu-rep --page details > $out/ncu_full_details_c2_jsd.txt 2>/dev/null
rm -f $out/prof_jsd_c2.ncu-rep
( timeout 200 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 ) > $out/bench_c2_reference_arm.json
grep -c tile_kernel $out/ncu_launches_bench_c2.csv
