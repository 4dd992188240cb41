#!/bin/bash
# column-blocked SpMV: tests + sweep of block counts on the C2 matrix (random 10M x 10M, 50/row)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_colblock.py -x -q 2>&1 | tail -5
for nb in 0 2 3 4; do SWEEP_COLBLOCK=$nb timeout 120 ./tools/spmv_sweep 10000000 50 10 random single 0 0 2 | tail -3; done
SWEEP_COLBLOCK=2 B2S_SPMV_TILE_NNZ=1024 timeout 120 ./tools/spmv_sweep 10000000 50 10 random single 0 1024 2 | tail -2
for nb in 0 4 6; do SWEEP_COLBLOCK=$nb timeout 120 ./tools/spmv_sweep 20000000 50 10 random single 0 0 2 | tail -3; done
SWEEP_COLBLOCK=0 timeout 120 ./tools/spmv_sweep 10000000 51 10 banded single 0 0 2 | tail -3
./tools/spmv_sweep 10000000 50 10 random single 0 0 2 | tail -1
./tools/spmv_sweep 10000000 51 10 banded single 0 0 2 | tail -1
