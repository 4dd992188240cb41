#!/bin/bash
# resident CTAs x ring depth x shared-memory carve-out of the products consumer (2 groups, 1024-nnz tiles)
mkdir -p gpurun_out
{
for cfg in "3 3 57" "3 3 80" "3 2 44" "3 4 80"; do
  set -- $cfg
  echo "=== stages=$1 ctas=$2 carveout=$3"
  export B2S_SPMV_STAGES=$1 B2S_SPMV_CTAS=$2 B2S_SPMV_CARVEOUT=$3
  echo -n "random C2:  "; SWEEP_COLBLOCK=0 timeout 120 tools/spmv_sweep 10000000 50 10 random | grep "colblock  " | cut -c50-120
  echo -n "poisson:    "; B2S_SPMV_TILE_NNZ=1024 timeout 120 tools/spmv_sweep 4096 5 20 poisson single 3 1024 2 | tail -1 | cut -c60-130
  echo -n "powerlaw:   "; python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c130-200
done
} > gpurun_out/r2_occ_sweep2.txt 2>&1
cat gpurun_out/r2_occ_sweep2.txt
