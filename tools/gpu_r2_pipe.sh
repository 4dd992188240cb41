#!/bin/bash
# round 2: products-consumer configurations of the pipe kernel on the column-blocked C2 matrix
mkdir -p gpurun_out
{
for xf in 0 2 4 6 8 10 16; do for cfg in "1 2048" "2 1024"; do
  set -- $cfg
  echo "xflags=$xf groups=$1 tile=$2"
  B2S_SPMV_XFLAGS=$xf SWEEP_COLBLOCK=0 B2S_SPMV_GROUPS=$1 B2S_SPMV_TILE_NNZ=$2 timeout 120 tools/spmv_sweep 10000000 50 10 random | grep "colblock  "
done; done
} > gpurun_out/r2_pipe_sweep.txt 2>&1
cat gpurun_out/r2_pipe_sweep.txt
B2S_SPMV_XFLAGS=0 SWEEP_COLBLOCK=0 B2S_SPMV_GROUPS=2 B2S_SPMV_TILE_NNZ=1024 timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmv_pipe -s 6 -c 2 -o gpurun_out/r2_pipe_g2 -f tools/spmv_sweep 10000000 50 3 random > gpurun_out/r2_ncu.log 2>&1
tail -3 gpurun_out/r2_ncu.log
