#!/bin/bash
# Debug aid for gpurun: run each GPU test group in its own process (a trapped kernel poisons only its group).
# usage: bash tests/gpu_probe.sh [pytest -k expressions...]
mkdir -p gpurun_out
groups=("$@")
if [ ${#groups[@]} -eq 0 ]; then
  groups=("gemm_f32" "gemm_16 or fp16_conv" "conv3x3_f32 or up2_parity" "conv3x3_16" "attention_f32" "attention_16" "layernorm or stft")
fi
i=0
for g in "${groups[@]}"; do
  i=$((i+1))
  echo "=== group $i: $g"
  timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$g" 2>&1 | tail -25 | tee gpurun_out/probe_$i.log
done
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
