#!/bin/bash
# full single-GPU validation: every GPU test, smoke, bench
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 ) > gpurun_out/r2_pytest_gpu.log
cat gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
tail -3 gpurun_out/r2_bench_n1.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_n1.json").read().strip().splitlines()[-1])
for k in ["value","ms_per_step"]: print(k, d[k])
print("roofline", d["roofline"]["frac"], d["roofline"]["l2_request_ceiling"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["max_abs_diff_vs_device_path"])
for k in ["cg","powerlaw","spgemm"]: print(k, json.dumps(d.get(k))[:700])
PY
