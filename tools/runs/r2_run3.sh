#!/bin/bash
# round-2 GPU call 3: new default bench line + ncu source-level capture of the glue kernels
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err ) 2> gpurun_out/r2c_time.txt
tail -3 gpurun_out/r2c_bench.err; cat gpurun_out/r2c_time.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"enhance_istft|stft_kernel|head_rows|stem_tc_kernel|skip_sample8" -c 8 -f -o gpurun_out/r2c_glue python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2c_ncu.log 2>&1
tail -3 gpurun_out/r2c_ncu.log
ls -la gpurun_out/
python - <<PY
import json
l=json.load(open('gpurun_out/r2c_bench.json'))
print(json.dumps({k:l[k] for k in ('value','ms_per_step','e2e','clocks','sustained','roofline','model_roofline','configs','cpu_baseline','latency')}, indent=1)[:6000])
for r in l['shapes']: print(r)
PY
