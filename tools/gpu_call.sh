#!/bin/bash
# One parametrised GPU-box script (replaces the per-round gpu_rNN.sh files): every section writes under gpurun_out/<tag>/.
#   tools/gpu_call.sh <tag> [sections...]     sections: tests tests:<file> smoke bench bench:<workload> refarm ncu_list[:<workload>]
#                                             ncu_full:<workload>:<demangled-name regex> sh:<script> py:<script>
tag=${1:-r23}; shift
out=gpurun_out/$tag
mkdir -p $out
for sec in "$@"; do
  case $sec in
    tests)    ( timeout 900 python -m pytest tests -m gpu -q -rfxXs --tb=short 2>&1 | tail -80 ) > $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log ;;
    tests:*)  f=${sec#tests:}; ( timeout 900 python -m pytest $f -m gpu -q -rfxXs -s --tb=short 2>&1 | tail -150 ) > $out/pytest_$(basename $f .py).log; tail -5 $out/pytest_$(basename $f .py).log ;;
    smoke)    ( timeout 200 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 ) > $out/smoke.log; cat $out/smoke.log ;;
    bench)    ( timeout 600 python bench.py 2>$out/bench.err | tail -1 ) > $out/bench_c2.json; cut -c1-600 $out/bench_c2.json ;;
    bench:*)  w=${sec#bench:}; ( timeout 600 python bench.py --workload $w 2>$out/bench_$w.err | tail -1 ) > $out/bench_$w.json; cut -c1-400 $out/bench_$w.json ;;
    refarm)   ( timeout 600 python bench.py --impl reference 2>$out/bench_ref.err | tail -1 ) > $out/bench_reference.json; cut -c1-400 $out/bench_reference.json ;;
    ncu_list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/ncu_launches_bench_c2.csv python bench.py --steps 12 --warmup 6 --no-extras --no-cpu-baseline --e2e-steps 3 > $out/ncu_list_bench.log 2>&1; grep -c dct $out/ncu_launches_bench_c2.csv ;;
    ncu_list:*) w=${sec#ncu_list:}; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $out/ncu_launches_bench_$w.csv python bench.py --workload $w --steps 8 --warmup 4 --no-extras --no-cpu-baseline --e2e-steps 3 > $out/ncu_list_bench_$w.log 2>&1; grep -c dct $out/ncu_launches_bench_$w.csv ;;
    ncu_full:*) spec=${sec#ncu_full:}; w=${spec%%:*}; k=${spec#*:}; timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$k -s 6 -c 2 -o $out/ncu_full_${w}_$(echo $k | tr -cd 'A-Za-z0-9_') -f python bench.py --workload $w --steps 8 --warmup 4 --no-extras --no-cpu-baseline --e2e-steps 3 > $out/ncu_full_${w}.log 2>&1; ls -la $out/*.ncu-rep ;;
    sh:*)     s=${sec#sh:}; ( timeout 1500 bash $s $out 2>&1 | tail -80 ) > $out/$(basename $s .sh).log; tail -40 $out/$(basename $s .sh).log ;;
    py:*)     s=${sec#py:}; ( timeout 1500 python $s 2>&1 | tail -80 ) > $out/$(basename $s .py).log; tail -40 $out/$(basename $s .py).log ;;
    *) echo "unknown section $sec" ;;
  esac
done
