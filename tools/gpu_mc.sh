#!/bin/bash
N=${1:-2}
( timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) 2>&1 | tail -3
for mc in 1 0; do
  LEGATE_SPARSE_MULTICAST=$mc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$mc tools/side_bench.py cg --grid 4096 --iters 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('multicast=$mc N=$N cg it/s', round(d['fused']['iters_per_s'],1), d['solve_rtol_1e-10']['iters'])"
  LEGATE_SPARSE_MULTICAST=$mc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2957$mc bench.py --gpus $N --steps 20 --warmup 5 --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('multicast=$mc N=$N spmv', round(d['value'],1), 'gathered', round(d['gathered']['value'],1), d['gathered']['ms_per_step'])"
done
