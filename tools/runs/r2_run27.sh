#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x 2>&1 | tail -3
for f in 0 1 0 1 0 1; do
  HVIT_SIDE_STREAM=$f python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2ae_bench_s$f.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2ae_bench_s$f.json')); print('side=$f', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), d['latency']['p50_ms'])
P
done
