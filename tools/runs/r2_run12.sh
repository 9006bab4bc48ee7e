#!/bin/bash
mkdir -p gpurun_out
A="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
show() { python - "$1" "$2" <<PY
import json,sys
l=json.load(open(sys.argv[1]))
d={r['name']:r['us_per_launch'] for r in l['shapes']}
print(sys.argv[2], round(l['value']), 'ms', round(l['ms_per_step'],4), 'patch', round(d['patch_embed'],1), 'fc2', round(d['blocks.*.fc2'],1), 'enc1', round(d['encoder.1'],1), 'enc2', round(d['encoder.2'],1))
PY
}
for PF in 0 2 4 8 12; do HVIT_A_PREFETCH=$PF python bench.py $A > gpurun_out/h_pf$PF.json 2>/dev/null; show gpurun_out/h_pf$PF.json a_prefetch=$PF; done
python bench.py $A > gpurun_out/h_def.json 2>/dev/null; show gpurun_out/h_def.json default
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "patch_embed or gemm_16" 2>&1 | tail -2
