#!/bin/bash
# round-2 GPU call 2: full GPU suite after the boundary / glue refactor + bench
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q -x > gpurun_out/r2b_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2b_tests.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2b_steps.json > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1
tail -15 gpurun_out/r2b_tests.log
tail -3 gpurun_out/r2b_smoke.log
tail -3 gpurun_out/r2b_bench.err
python - <<PY
import json
l=json.load(open('gpurun_out/r2b_bench.json'))
print(round(l['value']), round(l['e2e']['value']), l['ms_per_step'], l['clocks'], l['latency'])
d=json.load(open('gpurun_out/r2b_steps.json'))
for k,v in sorted(d['families'].items(), key=lambda kv:-kv[1]['ms']): print(k, round(v['ms']*1e3,1), v['launches'])
PY
