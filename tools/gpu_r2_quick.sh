#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmv.py tests/test_gpu_colblock.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do echo -n "random C2:  "; SWEEP_COLBLOCK=0 timeout 120 tools/spmv_sweep 10000000 50 10 random | grep "colblock  " | cut -c50-120; done
echo -n "poisson:    "; timeout 120 tools/spmv_sweep 4096 5 20 poisson single 3 1024 2 | tail -1 | cut -c60-130
echo -n "powerlaw:   "; python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c130-330
