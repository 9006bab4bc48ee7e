#!/bin/bash
mkdir -p gpurun_out
A="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
show() { python - "$1" "$2" <<PY
import json,sys
l=json.load(open(sys.argv[1]))
d={r['name']:r['us_per_launch'] for r in l['shapes']}
print(sys.argv[2], round(l['value']), 'ms', round(l['ms_per_step'],4), 'ln', round(d['blocks.*.norm1'],1), round(d['blocks.*.norm2'],1), 'head', round(d['decoder.3'],1), 'attn', round(d['blocks.*.attn'],1))
PY
}
for LB in 2 3 4 6; do HVIT_LN_BLOCKS=$LB python bench.py $A > gpurun_out/g_ln$LB.json 2>/dev/null; show gpurun_out/g_ln$LB.json ln_blocks=$LB; done
HVIT_HEAD_MINB=2 python bench.py $A > gpurun_out/g_head2.json 2>/dev/null; show gpurun_out/g_head2.json head_minb2
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "layernorm or head" 2>&1 | tail -2
