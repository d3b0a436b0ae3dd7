#!/bin/bash
# 1-GPU call r03: full GPU parity suite, smoke, A/B of the publish hook (same box), bench (both arms, c1..c4),
# c5 sweep at 64Mi / 256Mi / 1Gi pixels, ncu launch list + full capture of the c4 (C=19) JSD kernel.
tag=${1:-r03}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
nproc > $out/nproc.txt
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
for rep in 1 2; do
  for lib in product nopub; do
    if [ $lib = nopub ]; then export DCT_B200_LIB=$PWD/tools/ab/libdct_nopub.so; else unset DCT_B200_LIB; fi
    timeout 200 python bench.py --steps 3000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$lib rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f' % (d['ms_per_step'], r['kernel_ms']*1e3))"
  done
done > $out/ab_publish_hook.log 2>&1
unset DCT_B200_LIB
( timeout 600 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 ) > $out/bench_c2_reference_arm.json
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.json
( timeout 300 python bench.py --workload c1 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c1.json
( timeout 300 python bench.py --workload c3 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c3.json
( timeout 400 python tools/sweep.py --no-aten --reps 5 --out $out/sweep_64Mi 2>&1 | tail -12 ) > $out/sweep_64Mi.log
( timeout 400 python tools/sweep.py --no-aten --reps 3 --pixels 268435456 --ks 2,3 --cs 2,4,19 --mem-gb 40 --out $out/sweep_256Mi 2>&1 | tail -8 ) > $out/sweep_256Mi.log
( timeout 400 python tools/sweep.py --no-aten --reps 2 --pixels 1073741824 --ks 2,3 --cs 4,19 --mem-gb 60 --out $out/sweep_1Gi 2>&1 | tail -6 ) > $out/sweep_1Gi.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_bench_c4.csv \
    python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:JsdOp -s 4 -c 1 -o $out/prof_jsd_c4 \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_c4.log 2>&1
ncu -i $out/prof_jsd_c4.ncu-rep --page raw --csv > $out/ncu_full_raw_c4_jsd.csv 2>/dev/null
ncu -i $out/prof_jsd_c4.ncu-rep --page details > $out/ncu_full_details_c4_jsd.txt 2>/dev/null
tail -3 $out/pytest_gpu.log; cat $out/smoke.log $out/ab_publish_hook.log; cut -c1-400 $out/bench_c2_default.json; cut -c1-300 $out/bench_c4.json; cat $out/sweep_64Mi.log | cut -c1-200
