#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2g_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2g_tests.log
tail -5 gpurun_out/r2g_tests.log
bash tools/r2_sanitize.sh
