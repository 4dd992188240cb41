#!/bin/bash
N=${1:-4}
mkdir -p gpurun_out
for mc in 0 1; do
LEGATE_SPARSE_MULTICAST=$mc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$mc tools/gpu_r2_gathered.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM\|^$" | tail -3
done | tee gpurun_out/r2_gathered_n$N.txt
