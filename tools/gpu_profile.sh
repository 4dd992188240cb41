#!/bin/bash
# GPU validation + evidence pass: tests, bench, ncu launch list, one ncu --set full capture.
mkdir -p gpurun_out /tmp/ncu
( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
( timeout 300 python -c "import __graft_entry__ as g; g.smoke()" ) 2>&1 | tail -1
( timeout 900 python bench.py ) > gpurun_out/bench.log 2>gpurun_out/bench.err; echo "bench rc=$?"
( timeout 600 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"
# launch list (every kernel with its device time; cold-cache, serialised)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench.csv \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/launches_bench.log 2>&1
# top kernel, full set, inside the bench command
timeout 900 ncu --set full --clock-control none --import-source on -k regex:spmv_pipe -s 4 -c 2 -f -o /tmp/ncu/spmv_bench \
   python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page raw --csv > gpurun_out/ncu_spmv_bench_raw.csv 2>/dev/null
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page details --csv > gpurun_out/ncu_spmv_bench_details.csv 2>/dev/null
ncu -i /tmp/ncu/spmv_bench.ncu-rep --page source --csv > gpurun_out/ncu_spmv_bench_source.csv 2>/dev/null
cp /tmp/ncu/spmv_bench.ncu-rep gpurun_out/ 2>/dev/null
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/bench.log; tail -2 gpurun_out/bench.err; cat gpurun_out/bench_reference.log | tail -1
( SWEEP_COLBLOCK=0 ./tools/spmv_sweep 10000000 50 10 random single 0 0 2 | tail -2; ./tools/spmv_sweep 10000000 51 10 banded single 0 0 2 | tail -1 ) > gpurun_out/sweep_final.log 2>&1; cat gpurun_out/sweep_final.log
