#!/bin/bash
# multi-GPU pass (run with gpurun --gpus N): dist tests + bench at N ranks
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
( timeout 1200 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) > gpurun_out/pytest_dist_$N.log 2>&1; echo "dist pytest rc=$?"
tail -15 gpurun_out/pytest_dist_$N.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 tools/side_bench.py gmres --iters 60 ) 2>/dev/null | tail -1 | cut -c1-500 | tee gpurun_out/gmres_n$N.log
