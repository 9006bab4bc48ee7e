#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_model.py -m gpu -q -x -k "cli or device_metrics or evaluator or directory or c_host" > gpurun_out/r2h_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2h_tests.log
tail -30 gpurun_out/r2h_tests.log
