#!/bin/bash
mkdir -p gpurun_out
echo "=== powerlaw (single barrier long-row pass)"; python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-330
python -m pytest tests/test_gpu_spmv.py -m gpu -x -q 2>&1 | tail -2
# SpGEMM: launch list (per-kernel time) on R-MAT 18, then full captures of the two heaviest kernels
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_spgemm_launches.csv python tools/side_bench.py spgemm --scale 18 --verify-scale 0 --banded-n 1000000 > gpurun_out/r2_spgemm_side.log 2>&1
python tools/launch_summary.py gpurun_out/r2_spgemm_launches.csv 16
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"hash_kernel|dense_row" -c 14 -o gpurun_out/r2_spgemm -f python tools/side_bench.py spgemm --scale 16 --verify-scale 0 --banded-n 100000 > gpurun_out/r2_spgemm_ncu.log 2>&1
tail -2 gpurun_out/r2_spgemm_ncu.log
python tools/side_bench.py spgemm --scale 16 18 --verify-scale 16 --banded-n 4000000 2>/dev/null | cut -c1-400
