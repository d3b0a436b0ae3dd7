#!/bin/bash
# 1-GPU call r18: ncu --set full of the c2 normalisation cluster kernels (double pass, and single pass + clamp tail)
out=gpurun_out/${1:-r18}; mkdir -p $out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:l2_cluster -s 6 -c 2 -o $out/prof_l2_c2 \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 --graph 0 > $out/ncu_full_l2.log 2>&1
ncu -i $out/prof_l2_c2.ncu-rep --page raw --csv > $out/ncu_full_raw_c2_l2.csv 2>/dev/null
ncu -i $out/prof_l2_c2.ncu-rep --page details > $out/ncu_full_details_c2_l2.txt 2>/dev/null
ncu -i $out/prof_l2_c2.ncu-rep --page source --csv > $out/ncu_source_c2_l2.csv 2>/dev/null
rm -f $out/prof_l2_c2.ncu-rep
grep -n "Duration\|Stall\|stall\|Warp Cycles Per Issued\|Eligible\|No Eligible\|Active Warps\|DRAM Throughput\|Achieved Occupancy" $out/ncu_full_details_c2_l2.txt | head -40
