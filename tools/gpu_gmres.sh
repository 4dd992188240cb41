#!/bin/bash
# GMRES native kernels + CG timing investigation
timeout 900 python -m pytest tests/test_gpu_gmres_kernels.py tests/test_gpu_solvers.py -x -q 2>&1 | tail -4
timeout 300 python tools/side_bench.py gmres --grid 4096 --iters 100 2>/dev/null | tail -1
for it in 200 1000; do timeout 300 python tools/side_bench.py cg --grid 4096 --iters $it --no-solve 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('cg iters', d['iters'], 'fused it/s', round(d['fused']['iters_per_s'],1), 'unfused', round(d['unfused']['iters_per_s'],1))"; done
LEGATE_SPARSE_CG_GRAPH=0 timeout 300 python tools/side_bench.py cg --grid 4096 --iters 200 --no-solve 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nograph cg fused it/s', round(d['fused']['iters_per_s'],1))"
nvidia-smi --query-gpu=clocks.sm,clocks.mem,clocks.max.mem,power.draw,temperature.gpu,clocks_throttle_reasons.active --format=csv
