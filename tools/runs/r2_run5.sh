#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x -s -k "varlen or directory or evaluator" > gpurun_out/r2e_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2e_tests.log
grep -E "^\[varlen|passed|failed|rc=|Error|error" gpurun_out/r2e_tests.log | head -40
tail -30 gpurun_out/r2e_tests.log
