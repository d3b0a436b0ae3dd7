#!/bin/bash
# Round-2 experiment, prepared but NOT yet run on hardware: sub-tiles at the end of the tile pipeline's tail pool
# (csrc/dct_tile.cuh, DCT_POOL_SPLIT / DCT_POOL_SPLIT_LEVELS).
#
# Step 1 (here, no GPU): build the variant library and micro-benchmarks next to the product build
#   tools/ab_pool_split.sh build
# Step 2 (GPU box, one gpurun call): parity of the variant library first, then the A/B
#   gpurun --timeout 600 -- 'bash tools/ab_pool_split.sh run r23'
set -u
P=deep-co-training-for-semi-supervised-image-segmentation_b200/csrc
NV="nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo"
case "${1:-}" in
build)
  mkdir -p tools/ab
  for S in 2 4; do
    ( mkdir -p /tmp/ps$S && cd $P && for f in dct_abi dct_jsd dct_jsd_k2 dct_jsd_k3 dct_jsd_k4 dct_kl dct_ce dct_metrics dct_onehot dct_vat; do
        $NV -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DDCT_POOL_SPLIT=$S -c $f.cu -o /tmp/ps$S/$f.o & done; wait
      $NV -shared -o ../../tools/ab/libdct_b200_split$S.so /tmp/ps$S/*.o -cudart static ) &
    $NV -I$P -Iinclude -DDCT_POOL_SPLIT=$S tools/kbench_tile.cu -o tools/ab/kbench_split$S &
  done
  $NV -I$P -Iinclude tools/kbench_tile.cu -o tools/ab/kbench_split1 &
  wait; ls -la tools/ab ;;
run)
  out=gpurun_out/${2:-r23}; mkdir -p $out
  for S in 2 4; do   # the whole GPU parity suite against each variant library (DCT_B200_LIB: developer override, _lib.py)
    ( DCT_B200_LIB=$PWD/tools/ab/libdct_b200_split$S.so timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 ) > $out/pytest_gpu_split$S.log
    tail -2 $out/pytest_gpu_split$S.log
  done
  # kernels: jsd+dice c2 (0), klfromlogits (32), kllogit (40), Dice meter (49), jsd+dice c3 (57), c1 x 8 (64)
  for rep in 1 2; do for idx in 0 32 40 49 57 64; do for S in 1 2 4; do
    echo -n "split$S rep$rep "; timeout 60 tools/ab/kbench_split$S 30 $idx 32 1 0 2>&1 | grep -A1 "us " | tr '\n' ' '; echo
  done; done; done > $out/kbench_pool_split.log 2>&1
  cut -c1-260 $out/kbench_pool_split.log
  for wl in c2 c1 c3 c4; do for S in 1 2 4; do
    lib=$PWD/tools/ab/libdct_b200_split$S.so; [ $S = 1 ] && lib=
    DCT_B200_LIB=$lib timeout 200 python bench.py --workload $wl --steps 1000 --no-cpu-baseline --e2e-steps 5 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('split$S $wl ms_per_step=%.4f kernel_us=%.2f frac=%.3f' % (d['ms_per_step'], r['kernel_ms']*1e3, r['frac']))"
  done; done > $out/ab_step_pool_split.log 2>&1
  cat $out/ab_step_pool_split.log ;;
*) echo "usage: $0 build | run [tag]"; exit 2 ;;
esac
