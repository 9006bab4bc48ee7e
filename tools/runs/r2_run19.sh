#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm" 2>&1 | tail -15
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "forward_matches_oracle or headline or golden" 2>&1 | tail -15
for f in 0 1 0 1; do
  HVIT_LN_FOLD=$f python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2r_steps_f$f.json > gpurun_out/r2r_bench_f$f.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2r_bench_f$f.json')); print('fold=$f', d['value'], d['ms_per_step'], d['e2e']['value'])
P
done
python tools/steps.py gpurun_out/r2r_steps_f1.json | grep "blocks\|norm\|total"
