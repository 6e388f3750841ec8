#!/bin/bash
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; tail -4 gpurun_out/sanitize_$tool.log
done
