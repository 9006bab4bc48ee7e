#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -x > gpurun_out/r2k_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2k_tests.log
tail -4 gpurun_out/r2k_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/r2_profile.sh r2
