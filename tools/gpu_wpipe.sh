#!/bin/bash
( timeout 1200 python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -q -x ) 2>&1 | tail -4
for st in 2 3; do ./tools/spmv_sweep 10000000 50 10 random single 5 1024 $st | tail -1 | sed "s/single/wpipe-s$st/"; done
./tools/spmv_sweep 10000000 50 10 random single 0 0 2 | tail -1 | sed "s/single/AUTO-random/"
./tools/spmv_sweep 10000000 51 10 banded single 0 0 2 | tail -1 | sed "s/single/AUTO-banded/"
./tools/spmv_sweep 10000000 51 10 banded single 5 1024 2 | tail -1 | sed "s/single/wpipe-banded/"
(timeout 600 python tools/side_bench.py powerlaw) 2>/dev/null | tail -2 | head -1
(B2S_SPMV_VARIANT=wpipe timeout 600 python tools/side_bench.py powerlaw) 2>/dev/null | tail -2 | head -1
