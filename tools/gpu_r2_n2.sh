#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) 2>&1 | tail -5
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/side_bench.py gmres --iters 60 2>/dev/null | tail -1 | cut -c1-400
LEGATE_SPARSE_NO_HALO=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 tools/side_bench.py gmres --iters 60 2>/dev/null | tail -1 | cut -c1-400
LEGATE_SPARSE_MULTICAST=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 tools/gpu_r2_gathered.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -1
