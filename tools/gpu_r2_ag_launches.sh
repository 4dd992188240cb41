#!/bin/bash
# ncu launch list (device time of every kernel; cold-cache, serialised) of the config-5 side bench
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_powerlaw.csv \
   python tools/side_bench.py powerlaw > gpurun_out/r2_launches_powerlaw.log 2>&1
python tools/launch_summary.py gpurun_out/r2_launches_powerlaw.csv 12 | tee gpurun_out/r2_launch_summary_powerlaw.txt
