#!/bin/bash
mkdir -p gpurun_out
python - <<'P'
import torch
p=torch.cuda.get_device_properties(0)
print("L2", p.L2_cache_size/2**20, "MB")
import ctypes
cudart=ctypes.CDLL("libcudart.so.12")
v=ctypes.c_int()
for name,attr in (("MaxPersistingL2CacheSize",108),("MaxAccessPolicyWindowSize",109)):
    cudart.cudaDeviceGetAttribute(ctypes.byref(v), attr, 0); print(name, v.value/2**20, "MB")
P
run() {
  env "${@:2}" python bench.py --no-cpu-baseline --no-extras --profile-out gpurun_out/r2aa_steps_$1.json > gpurun_out/r2aa_bench_$1.json 2>/dev/null
  python - <<P
import json
d=json.load(open('gpurun_out/r2aa_bench_$1.json')); print('$1', round(d['value']), d['ms_per_step'], round(d['e2e']['value']))
P
  python tools/steps.py gpurun_out/r2aa_steps_$1.json | grep "norm1\|norm2\|blocks.\*.proj\|fc2"
}
run pin A=1
run nopin HVIT_NO_L2PIN=1
run pin2 A=1
run nopin2 HVIT_NO_L2PIN=1
