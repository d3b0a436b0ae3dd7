#!/bin/bash
out=${1:-gpurun_out/kb}; mkdir -p $out
( KB_DUMP=1 timeout 300 tools/kbench_c2 200 0 0 2>&1 ) > $out/kbench_c2_dump.log
( timeout 300 tools/kbench_c2 200 -1 0 2>&1 ) > $out/kbench_c2.log
( timeout 300 tools/kbench_c2 200 -1 1 2>&1 ) > $out/kbench_c3c1.log
grep -v "^ " $out/kbench_c2.log $out/kbench_c3c1.log
