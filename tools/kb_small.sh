#!/bin/bash
# c2-sized launch shapes: tile of 512 px / 8 warps / 2 CTAs per SM (product) vs 256 px / 4 warps / 3-4 CTAs per SM, row copies vs tensor maps
out=${1:-gpurun_out/kb}; mkdir -p $out
( timeout 600 tools/kbench_tile 200 -1 32 1 8 2>&1 ) | tee $out/kbench_small.log
