#!/bin/bash
mkdir -p gpurun_out
{
for a in 0 1; do echo "=== L1_ALLOC=$a"; B2S_SPMV_L1_ALLOC=$a timeout 300 tools/spmv_sweep 4096 5 20 poisson | grep -v "^device"; done
echo "=== auto"; timeout 300 tools/spmv_sweep 4096 5 20 poisson | grep "pipe groups"
} > gpurun_out/r2_poisson_sweep.txt 2>&1
cat gpurun_out/r2_poisson_sweep.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:spmv_pipe -s 4 -c 1 -o gpurun_out/r2_powerlaw -f python tools/side_bench.py powerlaw > gpurun_out/r2_ncu_pl.log 2>&1
tail -2 gpurun_out/r2_ncu_pl.log
python tools/side_bench.py cg --iters 1000 --no-solve 2>/dev/null | tail -1 | cut -c1-330
