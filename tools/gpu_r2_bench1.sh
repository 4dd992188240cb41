#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_spmv.py tests/test_gpu_solvers.py -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -5 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step"]: print(k, d[k])
print("roofline", d["roofline"]["frac"], d["roofline"]["l2_request_ceiling"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["max_abs_diff_vs_device_path"])
for k in ["banded","cg","powerlaw","spgemm","cusparse","cpu_baseline"]: print(k, json.dumps(d.get(k))[:1200])
PY
