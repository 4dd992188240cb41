#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) on the async-gather kernel's irregular-rows test
mkdir -p gpurun_out
timeout 80 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 \
  python -m pytest "tests/test_gpu_agather.py::test_agather_irregular_rows_types[float64-0]" -m gpu -x -q \
  > gpurun_out/r2_sanitizer_racecheck_agather.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|hazard|Error" gpurun_out/r2_sanitizer_racecheck_agather.log | head -12
