#!/bin/bash
# compute-sanitizer passes on small shapes (memcheck + racecheck + synccheck): pipe kernel (products
# groups, long rows, column blocks / accumulate), gallery kernels, SpGEMM classes, C ABI
mkdir -p gpurun_out
export PYTHONWARNINGS=ignore
( timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 20 \
    python -m pytest tests/test_gpu_spmv.py tests/test_gpu_cabi.py tests/test_gpu_spgemm.py tests/test_gpu_gallery.py tests/test_gpu_colblock.py -m gpu -q -x \
    -k "reference_shapes or ragged or irregular or longrows or golden or error_codes or unaligned or edge_helpers or known_answers or all_row_classes or variants_types or random_matches or dense_to_csr or dia_to_csr or colblock" ) > gpurun_out/sanitize_memcheck.log 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck.log | head -12
( timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 \
    python -m pytest tests/test_gpu_spmv.py tests/test_gpu_colblock.py -m gpu -q -x -k "ragged or irregular or longrows or golden or colblock" ) > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.log | head -8
( timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 --print-limit 10 \
    python -m pytest tests/test_gpu_spmv.py -m gpu -q -x -k "irregular or longrows" ) > gpurun_out/sanitize_synccheck.log 2>&1
echo "synccheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_synccheck.log | head -4
