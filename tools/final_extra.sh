#!/bin/bash
# extras of a validation call: BASELINE configs[4] sweep (no ATen arm), compute-sanitizer memcheck over the tile-pipeline tests,
# in-kernel step timeline.  Every command is bounded.
out=${1:-gpurun_out/x}; mkdir -p $out
( timeout 300 python tools/sweep.py --no-aten --out $out/sweep 2>&1 | tail -20 ) > $out/sweep.log; cat $out/sweep.log
( timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_supervised.py -m gpu -q -x -k "consistency_step or ragged or dice or fused or confusion or ce_" 2>&1 | tail -12 ) > $out/sanitizer_memcheck.log; tail -6 $out/sanitizer_memcheck.log
( timeout 120 python tools/step_trace.py 2>&1 | tail -8 ) > $out/step_trace_c2.log; cat $out/step_trace_c2.log
