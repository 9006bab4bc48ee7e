#!/bin/bash
# round-2 GPU call 1: new parity tests + MLP panel experiment
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv > gpurun_out/r2a_smi.txt
timeout 1500 python -m pytest tests/test_gpu_model.py -m gpu -q -x -s -k "v2 or literal or headline or widened or bf16_range or directory or evaluator" > gpurun_out/r2a_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/r2a_tests.log
for P in 1 2 4 8; do
  HVIT_MLP_PANELS=$P timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --profile-out gpurun_out/r2a_steps_p$P.json > gpurun_out/r2a_bench_p$P.json 2> gpurun_out/r2a_bench_p$P.err
done
tail -5 gpurun_out/r2a_tests.log
for P in 1 2 4 8; do python - <<PY
import json
l=json.load(open('gpurun_out/r2a_bench_p$P.json'))
print('panels $P', round(l['value']), round(l['e2e']['value']), l['ms_per_step'], l['clocks'])
PY
done
