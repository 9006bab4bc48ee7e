#!/bin/bash
# compute-sanitizer over the tiny-model smoke (memcheck / synccheck / racecheck / initcheck); summaries -> gpurun_out/
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_smoke.py fp16 > gpurun_out/r2_sanitizer_$tool.log 2>&1
  echo "exit=$?" >> gpurun_out/r2_sanitizer_$tool.log
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|exit=|sanitize smoke done|enhance max-rel|Error|hazard" gpurun_out/r2_sanitizer_$tool.log | head -12
done
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_smoke.py fp32 > gpurun_out/r2_sanitizer_memcheck_fp32.log 2>&1
echo "exit=$?" >> gpurun_out/r2_sanitizer_memcheck_fp32.log
grep -E "ERROR SUMMARY|exit=|sanitize smoke done" gpurun_out/r2_sanitizer_memcheck_fp32.log
