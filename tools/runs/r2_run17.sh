#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 4 5 7; do HVIT_DBG=$d timeout 120 python tests/gemm_probe.py 2>&1 | grep -v big; done > gpurun_out/r2p_gemm_dbg.log 2>&1
HVIT_PROF=1 timeout 120 python tests/gemm_probe.py 2>&1 | grep "igemm prof" | awk '!seen[$0]++' | head -12 >> gpurun_out/r2p_gemm_dbg.log
cat gpurun_out/r2p_gemm_dbg.log
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
