#!/bin/bash
# compute-sanitizer on the FINAL libtgs.so: bash profiles/sanitize.sh <tag>   (logs -> gpurun_out/<tag>_sanitizer_*.log)
TAG=$1
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python profiles/sanitize_case.py > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  tail -3 gpurun_out/${TAG}_sanitizer_${tool}.log
done
