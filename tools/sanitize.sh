#!/bin/bash
# compute-sanitizer over tools/sanitizer_case.py (GPU box): memcheck, racecheck, initcheck, synccheck.
# Summaries go to gpurun_out/sanitizer_<tool>.txt (copied to profiles/r02/ by hand).
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck synccheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check full"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 python tools/sanitizer_case.py > gpurun_out/sanitizer_$tool.full 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.full
  { echo "compute-sanitizer --tool $tool $extra python tools/sanitizer_case.py"; grep -E "sanitizer case done|ERROR SUMMARY|RACECHECK SUMMARY|LEAK SUMMARY|exit code|Error|error:" gpurun_out/sanitizer_$tool.full | head -40; } > gpurun_out/sanitizer_$tool.txt
  cat gpurun_out/sanitizer_$tool.txt
done
