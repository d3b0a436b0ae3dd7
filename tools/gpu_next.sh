#!/bin/bash
# First 1-GPU call of the next round (prepared at the end of round 1, when no GPU minutes were left):
#   1. the bf16 parity file that has never run on hardware (reported test by test: XPASS = verified, XFAIL = a bug to fix),
#   2. the whole GPU suite, smoke, the headline bench line,
#   3. if tools/ab/ holds the variant builds (tools/ab_pool_split.sh build), the tail-pool sub-tile A/B.
tag=${1:-r23}
out=gpurun_out/$tag
mkdir -p $out
( timeout 300 python -m pytest tests/test_zz_gpu_bf16.py -m gpu -q -rxX 2>&1 | tail -60 ) > $out/pytest_bf16.log
grep -c XPASS $out/pytest_bf16.log; grep XFAIL $out/pytest_bf16.log | cut -c1-200 | head -30
( timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 ) > $out/pytest_gpu.log
tail -2 $out/pytest_gpu.log
( timeout 120 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 ) > $out/smoke.log; cat $out/smoke.log
( timeout 400 python bench.py 2>&1 | tail -1 ) > $out/bench_c2_default.json; cut -c1-330 $out/bench_c2_default.json
[ -f tools/ab/libdct_b200_split2.so ] && bash tools/ab_pool_split.sh run $tag
