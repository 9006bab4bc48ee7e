#!/bin/bash
mkdir -p gpurun_out
for sms in 148 74 36; do for d in 0 5; do echo "sms=$sms"; HVIT_NUM_SMS=$sms HVIT_DBG=$d timeout 120 python tests/gemm_probe.py 2>&1 | grep "qkv\|fc2"; done; done > gpurun_out/r2q_gemm_sms.log 2>&1
cat gpurun_out/r2q_gemm_sms.log
