#!/bin/bash
# attention exp-phase ping-pong: A/B standalone timing, role profile, kernel tests, bench
mkdir -p gpurun_out
for pp in 0 1; do
  echo "== HVIT_ATTN_HALF=$pp"
  HVIT_ATTN_HALF=$pp timeout 300 python tests/attn_probe.py 2>&1 | grep -v "attn prof"
  HVIT_ATTN_HALF=$pp HVIT_PROF=1 timeout 300 python tests/attn_probe.py 2>&1 | grep "N=496\|N=1248" | awk 'NR%20==1'
done > gpurun_out/r2ag_attn.log 2>&1
cat gpurun_out/r2ag_attn.log
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" 2>&1 | tail -3
for pp in 0 1 0 1; do
  HVIT_ATTN_HALF=$pp python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2ag_steps_pp$pp.json > gpurun_out/r2ag_bench_pp$pp.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2ag_bench_pp$pp.json')); print('pp=$pp', d['value'], d['ms_per_step'], d['e2e']['value'])
P
  python tools/steps.py gpurun_out/r2ag_steps_pp$pp.json | grep attn
done
