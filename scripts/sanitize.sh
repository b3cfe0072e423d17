#!/bin/bash
# compute-sanitizer evidence: memcheck, racecheck, synccheck on the small cases of scripts/sanitize_case.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_case.py 8 > gpurun_out/sanitizer_$tool.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|max\||done|Error|hazard" gpurun_out/sanitizer_$tool.log | head -12
done
