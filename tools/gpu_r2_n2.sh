#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 tools/side_bench.py cg --iters 3000 --no-solve 2>/dev/null | tail -1 | cut -c150-330
LEGATE_SPARSE_CG_NO_FUSED_EXCHANGE=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tools/side_bench.py cg --iters 3000 --no-solve 2>/dev/null | tail -1 | cut -c150-330
