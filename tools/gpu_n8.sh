#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ) > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -1 gpurun_out/bench_n$N.log | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('N=$N spmv', round(d['value'],1), 'gathered', round(d['gathered']['value'],1), 'banded', round(d['banded']['value'],1), 'cg', round(d['cg']['iters_per_s'],1), 'e2e', round(d['e2e']['value'],1))"
grep -i "warn\|error" gpurun_out/bench_n$N.err | head -5
for mc in 1; do
  LEGATE_SPARSE_MULTICAST=$mc timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2956$mc tools/side_bench.py cg --grid 4096 --iters 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('multicast=$mc N=$N cg it/s', round(d['fused']['iters_per_s'],1), d['solve_rtol_1e-10']['iters'])"
done
LEGATE_SPARSE_CG_GRAPH=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29577 tools/side_bench.py cg --grid 4096 --iters 200 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('nograph N=$N cg it/s', round(d['fused']['iters_per_s'],1))"
