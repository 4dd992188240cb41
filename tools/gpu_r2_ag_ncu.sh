#!/bin/bash
# ncu --set full of the async-gather kernel on config 5 (power-law), exported to CSV on the box
mkdir -p gpurun_out /tmp/ncu
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_agather" -s 3 -c 1 \
   -f -o /tmp/ncu/ag python tools/side_bench.py powerlaw > gpurun_out/r2_ag_ncu.log 2>&1
tail -2 gpurun_out/r2_ag_ncu.log | cut -c1-300
ncu -i /tmp/ncu/ag.ncu-rep --page raw --csv > gpurun_out/r2_ag_raw.csv 2>/dev/null
ncu -i /tmp/ncu/ag.ncu-rep --page details --csv > gpurun_out/r2_ag_details.csv 2>/dev/null
ncu -i /tmp/ncu/ag.ncu-rep --page source --csv > gpurun_out/r2_ag_source.csv 2>/dev/null
ls -la gpurun_out/r2_ag_*
