#!/bin/bash
# 1-GPU call r04: GPU parity suite (adds fused CE+confusion, *_pub exchange kernels), A/B on the same box of
# {product, product + loopback exchange, build without any publication code} for c2 and c4, c4 bench, ncu captures.
tag=${1:-r04}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks.mem,power.limit --format=csv > $out/gpu.csv 2>&1
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -5 ) > $out/smoke.log
for wl in c2 c4; do
  steps=3000; [ $wl = c4 ] && steps=300
  for rep in 1 2; do
    for lib in product loopback nopub; do
      unset DCT_B200_LIB; ex=auto
      [ $lib = nopub ] && export DCT_B200_LIB=$PWD/tools/ab/libdct_nopub.so
      [ $lib = loopback ] && ex=p2p
      timeout 200 python bench.py --workload $wl --steps $steps --no-cpu-baseline --e2e-steps 5 --exchange $ex 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('$wl $lib rep$rep ms_per_step=%.4f jsd_kernel_us=%.2f frac=%.3f' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac']))"
    done
  done
done > $out/ab_publish.log 2>&1
unset DCT_B200_LIB
( timeout 600 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json
( timeout 300 python bench.py --workload c4 --steps 300 --no-cpu-baseline 2>&1 | tail -1 ) > $out/bench_c4.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/ncu_launches_bench_c4.csv \
    python bench.py --workload c4 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 3 > $out/ncu_launch_bench_c4.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:JsdOp -s 4 -c 1 -o $out/prof_jsd_c4 \
    python bench.py --workload c4 --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_c4.log 2>&1
ncu -i $out/prof_jsd_c4.ncu-rep --page raw --csv > $out/ncu_full_raw_c4_jsd.csv 2>/dev/null
ncu -i $out/prof_jsd_c4.ncu-rep --page details > $out/ncu_full_details_c4_jsd.txt 2>/dev/null
tail -3 $out/pytest_gpu.log; cat $out/smoke.log $out/ab_publish.log; cut -c1-300 $out/bench_c4.json; tail -3 $out/ncu_full_c4.log
