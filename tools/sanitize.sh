#!/bin/bash
# compute-sanitizer over the library on a small workload (smoke-sized forward + training step): logs under gpurun_out/
for tool in memcheck racecheck initcheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 python tools/sanitize_workload.py > gpurun_out/sanitize_$tool.log 2>&1
  tail -3 gpurun_out/sanitize_$tool.log
done
