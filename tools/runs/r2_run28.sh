#!/bin/bash
# final bench lines (no ncu): same commands as tools/r2_profile.sh
mkdir -p gpurun_out
T=r2
python bench.py --profile-out gpurun_out/${T}_bench_steps.json --latency-sweep gpurun_out/${T}_latency_sweep.json > gpurun_out/${T}_bench_line.json 2> gpurun_out/${T}_bench.err
python bench.py --precision bf16 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_bf16.json 2>> gpurun_out/${T}_bench.err
python bench.py --seconds 10 --batch 24 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_10s_b24.json 2>> gpurun_out/${T}_bench.err
python bench.py --seconds 1 --batch 128 --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_1s_b128.json 2>> gpurun_out/${T}_bench.err
python bench.py --widened --no-cpu-baseline --no-extras > gpurun_out/${T}_bench_line_widened.json 2>> gpurun_out/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_reference_line.json 2>> gpurun_out/${T}_bench.err
tail -3 gpurun_out/${T}_bench.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_line.json'))
print('value',round(d['value']),'ms',d['ms_per_step'],'e2e',round(d['e2e']['value']),'lat',d['latency']['p50_ms'], 'frac', d['roofline']['frac'], 'model', d['model_roofline']['frac_of_burst_bf16_peak'], 'sustained', round(d['sustained']['value']))
for f in ('bf16','10s_b24','1s_b128','widened'):
    x=json.load(open(f'gpurun_out/r2_bench_line_{f}.json')); print(f, round(x['value']), x['ms_per_step'])
P
