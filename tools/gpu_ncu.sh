#!/bin/bash
# ncu captures of single SpMV configurations (args: list of "tag:variant:tile:stages:mode[:i64]")
# Reports are exported to CSV on the box (gpurun_out is capped at 64 MiB); only the first .ncu-rep is kept.
mkdir -p gpurun_out /tmp/ncu
first=1
for spec in "$@"; do
  IFS=: read tag variant tile stages mode i64 <<< "$spec"
  k=50; [ "$mode" = "banded" ] && k=51
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spmv_(pipe|tile|rowvec)" -s 2 -c 1 \
     -f -o /tmp/ncu/ncu_$tag ./tools/spmv_sweep 10000000 $k 1 $mode single $variant $tile $stages $i64 > gpurun_out/ncu_$tag.log 2>&1
  tail -1 gpurun_out/ncu_$tag.log
  ncu -i /tmp/ncu/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
  ncu -i /tmp/ncu/ncu_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
  ncu -i /tmp/ncu/ncu_$tag.ncu-rep --page source --csv > gpurun_out/ncu_${tag}_source.csv 2>/dev/null
  if [ $first = 1 ]; then cp /tmp/ncu/ncu_$tag.ncu-rep gpurun_out/; first=0; fi
done
ls -la gpurun_out/ | head -40
