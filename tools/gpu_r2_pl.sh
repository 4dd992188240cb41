#!/bin/bash
mkdir -p gpurun_out
{
echo "=== powerlaw default"; python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== powerlaw colblock=2"; B2S_SPMV_COLBLOCK=2 python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== powerlaw colblock=3"; B2S_SPMV_COLBLOCK=3 python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== powerlaw variant=tile"; B2S_SPMV_VARIANT=tile python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== powerlaw variant=rowvec"; B2S_SPMV_VARIANT=rowvec python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== powerlaw 3 CTAs"; B2S_SPMV_CTAS=3 B2S_SPMV_CARVEOUT=80 python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
echo "=== poisson pipe 3 CTAs/SM default carveout"; B2S_SPMV_CTAS=3 B2S_SPMV_CARVEOUT=80 timeout 300 tools/spmv_sweep 4096 5 20 poisson | grep "pipe groups"
echo "=== poisson pipe 4 CTAs/SM"; B2S_SPMV_CTAS=4 B2S_SPMV_CARVEOUT=100 timeout 300 tools/spmv_sweep 4096 5 20 poisson | grep "pipe groups"
} > gpurun_out/r2_pl_variants.txt 2>&1
cat gpurun_out/r2_pl_variants.txt
