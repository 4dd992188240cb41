#!/bin/bash
# GPU validation + evidence pass (round 2): tests, smoke, bench (both arms), ncu launch list, one
# ncu --set full capture of the dominant kernel inside the bench command, exports for profiles/.
mkdir -p gpurun_out /tmp/ncu
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
( timeout 900 python bench.py ) > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"
# launch list (every kernel with its device time; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/launches_bench.csv 16 > gpurun_out/launch_summary.txt
# top kernel, full set, inside the bench command: the two launches of one timed step
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe -s 6 -c 2 -f -o /tmp/ncu/spmv_bench \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page raw --csv > gpurun_out/ncu_spmv_bench_raw.csv 2>/dev/null
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page details --csv > gpurun_out/ncu_spmv_bench_details.csv 2>/dev/null
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page source --csv > gpurun_out/ncu_spmv_bench_source.csv 2>/dev/null
python tools/ncu_top.py gpurun_out/ncu_spmv_bench_source.csv > gpurun_out/ncu_spmv_bench_source_top.txt 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; tail -c 1500 gpurun_out/bench.log; tail -2 gpurun_out/bench.err; tail -c 900 gpurun_out/bench_reference.log
( SWEEP_COLBLOCK=0 ./tools/spmv_sweep 10000000 50 10 random | tail -2; ./tools/spmv_sweep 10000000 51 10 banded | grep "pipe " ) > gpurun_out/sweep_final.log 2>&1; cat gpurun_out/sweep_final.log
cat gpurun_out/launch_summary.txt | head -8
