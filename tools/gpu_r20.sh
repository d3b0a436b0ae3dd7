#!/bin/bash
# 1-GPU call r20: Dice counters in the consumers' registers, handed to the producer warp on MARKED tiles only
# (DCT_DICE_LOCAL=1, csrc/dct_tile.cuh) against the per-tile fold (=0, tools/ab/libdct_b200_fold.so, tools/kbench_dice_fold):
# parity suite on the new build, kernel A/B, step A/B at c2 / c1 / c3, headline bench line, ncu launch list.
tag=${1:-r20}
out=gpurun_out/$tag
mkdir -p $out
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > $out/pytest_gpu.log
tail -3 $out/pytest_gpu.log
for rep in 1 2; do for idx in 0 49 57 64; do for v in fold local; do
  echo -n "$v rep$rep "; timeout 60 tools/kbench_dice_$v 30 $idx 32 1 0 2>&1 | grep -v trace
done; done; done > $out/kbench_dice_ab.log 2>&1
cat $out/kbench_dice_ab.log | cut -c1-150
step() {  # $1 = label, $2 = workload, env DCT_B200_LIB
  timeout 200 python bench.py --workload $2 --steps 2000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$1 $2 ms_per_step=%.4f kernel_us=%.2f frac=%.3f' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac']))"
}
for wl in c2 c1 c3 c2 c1 c3; do
  DCT_B200_LIB=$PWD/tools/ab/libdct_b200_fold.so step fold $wl
  step local $wl
done > $out/ab_step.log 2>&1
cat $out/ab_step.log
( timeout 400 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json
cut -c1-400 $out/bench_c2_default.json
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
cat $out/smoke.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/ncu_launches_bench_c2.csv \
    python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c2.log 2>&1
( timeout 200 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
( timeout 200 python bench.py --workload c1 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c1.json
grep -c tile_kernel $out/ncu_launches_bench_c2.csv
