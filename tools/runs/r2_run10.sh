#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_n2.json 2> gpurun_out/r2i_bench_n2.err
echo "rc=$?"; tail -3 gpurun_out/r2i_bench_n2.err
python - <<PY
import json
l=json.loads(open('gpurun_out/r2i_bench_n2.json').read().strip().splitlines()[-1])
print({k:l[k] for k in ('value','n_gpus','ms_per_step','clocks')}, l['e2e'], l['config'].get('host_cores_per_rank'))
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2i_ref_n2.json 2> gpurun_out/r2i_ref_n2.err
echo "rc=$?"; cut -c1-400 gpurun_out/r2i_ref_n2.json
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q -x -k "poison or second_device or edge_cases" 2>&1 | tail -3
