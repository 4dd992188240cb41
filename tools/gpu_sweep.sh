#!/bin/bash
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
( timeout 300 ./tools/spmv_sweep 10000000 50 10 random ) > gpurun_out/sweep_random.log 2>&1
( timeout 300 ./tools/spmv_sweep 10000000 51 10 banded ) > gpurun_out/sweep_banded.log 2>&1
tail -40 gpurun_out/pytest_gpu.log; cat gpurun_out/sweep_random.log gpurun_out/sweep_banded.log
