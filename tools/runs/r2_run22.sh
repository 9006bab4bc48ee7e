#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  env "${@:2}" python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2u_steps_$1.json > gpurun_out/r2u_bench_$1.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2u_bench_$1.json')); print('$1', round(d['value']), d['ms_per_step'])
P
  python tools/steps.py gpurun_out/r2u_steps_$1.json | grep "patch_embed\|fc2"
}
run base A=1
run patch128 HVIT_PATCH_BN=128
run fc2_128 HVIT_FC2_BN=128
run base2 A=1
run both HVIT_PATCH_BN=128 HVIT_FC2_BN=128
