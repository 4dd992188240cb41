#!/bin/bash
mkdir -p gpurun_out
for mode in random banded; do
  k=50; [ "$mode" = "banded" ] && k=51
  for tile in 1024 2048; do
    for v in 3 4 2; do
      ./tools/spmv_sweep 10000000 $k 10 $mode single $v $tile 2 | tail -1
    done
    B2S_SPMV_ROWWALK=1 ./tools/spmv_sweep 10000000 $k 10 $mode single 3 $tile 2 | tail -1 | sed 's/single/rowwalk/'
    B2S_SPMV_PRODUCTS=1 ./tools/spmv_sweep 10000000 $k 10 $mode single 3 $tile 2 | tail -1 | sed 's/single/product/'
  done
done 2>&1 | tee gpurun_out/sweep2.log
