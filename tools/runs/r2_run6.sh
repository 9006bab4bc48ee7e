#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "attention" > gpurun_out/r2f_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2f_tests.log
tail -4 gpurun_out/r2f_tests.log
timeout 300 python tests/attn_probe.py 2>&1 | grep "^B=" 
HVIT_PROF=1 timeout 300 python tests/attn_probe.py 2>&1 | grep "prof B=64 N=496" | head -2
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --profile-out gpurun_out/r2f_steps.json > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err
tail -2 gpurun_out/r2f_bench.err
python - <<PY
import json
l=json.load(open('gpurun_out/r2f_bench.json'))
print(round(l['value']), round(l['e2e']['value']), l['ms_per_step'], l['clocks'])
for r in l['shapes']:
    if r['launches']: print(r['name'], round(r['us_per_launch'],1))
PY
