#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "layernorm or folded or gemm_16" 2>&1 | tail -4
for d in 0 32; do HVIT_DBG=$d timeout 120 python tests/gemm_probe.py 2>&1 | grep "proj\|fc2"; done
for f in 0 1 0 1; do
  HVIT_LN_FOLD=$f python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2w_steps_f$f.json > gpurun_out/r2w_bench_f$f.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2w_bench_f$f.json')); print('fold=$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
P
done
python tools/steps.py gpurun_out/r2w_steps_f1.json | grep "blocks\|norm\|to_feature\|total"
python tools/plan_latency.py 2>&1 | tail -8
