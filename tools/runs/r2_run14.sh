#!/bin/bash
mkdir -p gpurun_out
A="--steps 20 --warmup 5 --no-cpu-baseline --no-extras"
show() { python - "$1" "$2" <<PY
import json,sys
l=json.load(open(sys.argv[1]))
d={r['name']:r['us_per_launch'] for r in l['shapes']}
print(sys.argv[2], round(l['value']), 'ms', round(l['ms_per_step'],4), 'istft', round(d['istft'],1), 'stft', round(d['stft']*2,1), 'head', round(d['decoder.3'],1))
PY
}
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "golden or varlen or edge or poison" 2>&1 | tail -2
for i in 1 2; do
HVIT_ISTFT_FR=8 python bench.py $A > gpurun_out/j_a$i.json 2>/dev/null; show gpurun_out/j_a$i.json fr8_$i
python bench.py $A > gpurun_out/j_b$i.json 2>/dev/null; show gpurun_out/j_b$i.json fr16_$i
done
