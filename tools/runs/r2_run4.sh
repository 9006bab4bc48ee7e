#!/bin/bash
# round-2 GPU call 4: ncu source-level capture of attention + one K=512 GEMM
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"attn_tc_kernel" -s 6 -c 1 -f -o gpurun_out/r2d_attn python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r2d_ncu.log 2>&1
tail -2 gpurun_out/r2d_ncu.log
HVIT_PROF=1 timeout 300 python tests/attn_probe.py > gpurun_out/r2d_attn_prof.log 2>&1
tail -5 gpurun_out/r2d_attn_prof.log
ls -la gpurun_out/*.ncu-rep
