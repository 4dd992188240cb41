#!/bin/bash
# async-gather kernel: parity tests, config 5 with the kernel on / off, resident CTAs
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_agather.py -m gpu -x -q 2>&1 | tail -15 | tee /tmp/ag_t.log
grep -q "passed" /tmp/ag_t.log && ! grep -q "failed" /tmp/ag_t.log || exit 1
(
for cfg in "AGATHER=1" "AGATHER=1 B2S_SPMV_CTAS=3"; do
  echo "=== powerlaw B2S_SPMV_$cfg"
  env B2S_SPMV_$cfg timeout 40 python tools/side_bench.py powerlaw 2>/dev/null | head -1 | cut -c1-420
done
) > gpurun_out/r2_ag_powerlaw_exp.txt 2>&1
cat gpurun_out/r2_ag_powerlaw_exp.txt
