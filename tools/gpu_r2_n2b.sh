#!/bin/bash
# 2-GPU pass after the async-gather kernel: dist tests + bench at N=2 (run with gpurun --gpus 2)
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dist.py -m gpu -q -x ) > gpurun_out/r2b_pytest_dist_2.log 2>&1; echo "dist pytest rc=$?"
tail -4 gpurun_out/r2b_pytest_dist_2.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/r2b_bench_n2.json 2> gpurun_out/r2b_bench_n2.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2b_bench_n2.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"])
for k in ["gathered","cg","powerlaw","spgemm"]: print(k, json.dumps(d.get(k))[:500])
PY
