#!/bin/bash
# One parametrised GPU-box script (replaces the per-round gpu_rNN.sh files): every section writes under gpurun_out/<tag>/.
#   tools/gpu_call.sh <tag> [sections...]     sections: tests smoke bench ncu_list ncu_full:<regex> ab:<script> py:<script>
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
    ncu_list) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/ncu_launches_bench_c2.csv python bench.py --steps 2 --warmup 1 > $out/ncu_list_bench.log 2>&1; grep -c dct $out/ncu_launches_bench_c2.csv ;;
    ncu_full:*) k=${sec#ncu_full:}; timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -c 2 -o $out/ncu_full_$k -f python bench.py --steps 2 --warmup 1 > $out/ncu_full_$k.log 2>&1; ls -la $out/ncu_full_$k.ncu-rep ;;
    sh:*)     s=${sec#sh:}; ( timeout 1500 bash $s $out 2>&1 | tail -80 ) > $out/$(basename $s .sh).log; tail -40 $out/$(basename $s .sh).log ;;
    py:*)     s=${sec#py:}; ( timeout 1500 python $s 2>&1 | tail -80 ) > $out/$(basename $s .py).log; tail -40 $out/$(basename $s .py).log ;;
    *) echo "unknown section $sec" ;;
  esac
done
