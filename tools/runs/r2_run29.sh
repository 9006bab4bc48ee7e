#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "stem" 2>&1 | tail -6
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x -s -k "forward_matches_oracle or headline" 2>&1 | grep -E "per-stage|output max-rel|passed|failed|Error|assert" | cut -c1-400 | head -30
for i in 1 2 3; do
  python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2af_steps.json > gpurun_out/r2af_bench.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2af_bench.json')); print(round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['latency']['p50_ms'])
P
done
python tools/steps.py gpurun_out/r2af_steps.json | grep "encoder.0 \|encoder.1 \|total"
