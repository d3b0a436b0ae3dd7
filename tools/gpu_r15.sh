#!/bin/bash
# 1-GPU call r15: tail-pool size of the tile schedule (pool_div: the last 1/pool_div of every CTA's range is handed out
# dynamically; product default 5) at c2 sizes, product launch shapes only.
out=gpurun_out/${1:-r15}; mkdir -p $out
for pd in 5 1 2 3 8 16; do
  echo "== pool_div=$pd"
  timeout 200 tools/kbench_reg 30 -1 32 1 0 0 $pd 2>&1 | grep -A1 "thr= 288 stages=[4-7] minb=2" | grep -A1 "jsd+dice c2 \|klfromlogits\|kllogit\|jsd+dice c3\|jsd+dice c1x8" | grep -v "^--"
done > $out/pool_div.log 2>&1
cat $out/pool_div.log | cut -c1-200
