#!/bin/bash
mkdir -p gpurun_out
A="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
show() { python - "$1" "$2" <<PY
import json,sys
l=json.load(open(sys.argv[1]))
d={r['name']:r['us_per_launch'] for r in l['shapes']}
print(sys.argv[2], round(l['value']), 'ms', round(l['ms_per_step'],4), 'patch', round(d['patch_embed'],1), 'fc2', round(d['blocks.*.fc2'],1), 'fc1', round(d['blocks.*.fc1'],1), l['clocks']['sm_mhz'])
PY
}
for i in 1 2 3; do
python bench.py $A > gpurun_out/i_a$i.json 2>/dev/null; show gpurun_out/i_a$i.json off_$i
HVIT_A_PREFETCH=4 python bench.py $A > gpurun_out/i_b$i.json 2>/dev/null; show gpurun_out/i_b$i.json pf4_$i
done
