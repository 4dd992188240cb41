#!/bin/bash
N=${1:-1}
if [ "$N" = "1" ]; then
  ( timeout 900 python -m pytest tests/test_gpu_solvers.py tests/test_gpu_fullsize.py -m gpu -q -x ) 2>&1 | tail -3
  for g in 1 0; do LEGATE_SPARSE_CG_GRAPH=$g timeout 600 python tools/side_bench.py cg --grid 4096 --iters 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph=$g N=1', d['fused']['iters_per_s'], d['solve_rtol_1e-10'])"; done
  LEGATE_SPARSE_CG_GRAPH=1 timeout 600 python tools/side_bench.py cg --grid 1024 --iters 500 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph=1 grid1024', d['fused']['iters_per_s'])"
  LEGATE_SPARSE_CG_GRAPH=0 timeout 600 python tools/side_bench.py cg --grid 1024 --iters 500 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph=0 grid1024', d['fused']['iters_per_s'])"
else
  ( timeout 900 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) 2>&1 | tail -3
  for g in 2 0; do LEGATE_SPARSE_CG_GRAPH=$g timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2955$g tools/side_bench.py cg --grid 4096 --iters 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('graph=$g N=$N', d['fused']['iters_per_s'], d['solve_rtol_1e-10'])"; done
fi
