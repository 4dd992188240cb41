#!/bin/bash
# compute-sanitizer passes on small shapes (memcheck + racecheck)
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_spmv.py tests/test_gpu_cabi.py tests/test_gpu_spgemm.py -m gpu -q -x \
    -k "reference_shapes or ragged or wpipe or golden or error_codes or unaligned or edge_helpers or known_answers or all_row_classes or variants_types" ) > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck.log | head -12
( timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 \
    python -m pytest tests/test_gpu_spmv.py -m gpu -q -x -k "ragged or wpipe or golden" ) > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head -8
