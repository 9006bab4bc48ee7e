#!/bin/bash
mkdir -p gpurun_out
for f in 0 1 0 1 0 1; do
  HVIT_LN_FOLD=$f python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2y_steps_f$f.json > gpurun_out/r2y_bench_f$f.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2y_bench_f$f.json')); print('fold=$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
P
done
for f in 0 1; do python tools/steps.py gpurun_out/r2y_steps_f$f.json | grep "blocks\|transformer.norm\|to_feature\|total"; done
