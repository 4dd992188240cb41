#!/bin/bash
mkdir -p gpurun_out /tmp/ncu
( timeout 900 python -m pytest tests/test_gpu_examples.py -m gpu -q -x ) > gpurun_out/pytest_examples.log 2>&1; echo "examples rc=$?"; tail -15 gpurun_out/pytest_examples.log
( cd examples && timeout 300 python gmg.py -n 1024 -l 6 ) 2>&1 | tail -2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe -s 8 -c 1 -f -o /tmp/ncu/pl python tools/side_bench.py powerlaw > gpurun_out/ncu_pl.log 2>&1
ncu -i /tmp/ncu/pl.ncu-rep --page raw --csv > gpurun_out/ncu_pl_raw.csv 2>/dev/null
ncu -i /tmp/ncu/pl.ncu-rep --page source --csv > gpurun_out/ncu_pl_source.csv 2>/dev/null
ls -la gpurun_out/ncu_pl*
