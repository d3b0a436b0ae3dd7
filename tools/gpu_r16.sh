#!/bin/bash
# 1-GPU call r16: compute-sanitizer memcheck over a slice of the GPU parity suite (tile pipeline, wide JSD, fused CE +
# confusion, large-sample normalisation, exchange), then the full suite + smoke once more on the final build.
out=gpurun_out/${1:-r16}; mkdir -p $out
( timeout 800 compute-sanitizer --tool memcheck --error-exitcode 99 --print-limit 20 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_supervised.py tests/test_gpu_parity.py -m gpu -x -q \
    -k "exchange or loopback or confusion or wide or l2_normalize or consistency_step_vs_oracle" 2>&1 | tail -15 ) > $out/sanitizer_memcheck.log
( timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $out/pytest_gpu.log
( timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' 2>&1 | tail -3 ) > $out/smoke.log
cat $out/sanitizer_memcheck.log; tail -2 $out/pytest_gpu.log; cat $out/smoke.log
