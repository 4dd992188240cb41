#!/bin/bash
# A/B on one box: round-1 tree (.ab/r1) vs the current tree — CG iteration rate + per-kernel durations
mkdir -p gpurun_out
{
echo "=== r1 tree"; python .ab/r1/tools/side_bench.py cg --iters 1000 --no-solve 2>/dev/null | tail -1
echo "=== r2 tree"; python tools/side_bench.py cg --iters 1000 --no-solve 2>/dev/null | tail -1
echo "=== r2 tree, 1-iteration graphs only (LEGATE_SPARSE_CG_GRAPH_N=0)"; LEGATE_SPARSE_CG_GRAPH_N=0 python tools/side_bench.py cg --iters 1000 --no-solve 2>/dev/null | tail -1
} > gpurun_out/r2_cgab.txt 2>&1
LEGATE_SPARSE_CG_GRAPH=0 LEGATE_SPARSE_CG_UNFUSED=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cg_launches_r1.csv python .ab/r1/tools/side_bench.py cg --iters 30 --no-solve > /dev/null 2>&1
LEGATE_SPARSE_CG_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_cg_launches_r2.csv python tools/side_bench.py cg --iters 30 --no-solve > /dev/null 2>&1
cat gpurun_out/r2_cgab.txt
for f in r1 r2; do echo "--- $f"; python tools/launch_summary.py gpurun_out/r2_cg_launches_$f.csv 2>/dev/null | head -14; done
echo "=== sweep after fence removal"
SWEEP_COLBLOCK=0 timeout 120 tools/spmv_sweep 10000000 50 10 random | grep "colblock  "
echo "=== powerlaw longrows on/off"
for lr in 1 0; do B2S_SPMV_LONGROWS=$lr python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-400; done
