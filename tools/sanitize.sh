#!/bin/bash
# compute-sanitizer over every kernel family (tools/sanitize_cases.py): memcheck, racecheck, synccheck.
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh r02'
TAG=${1:-r02}
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
  echo "== $tool: exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|path |trunk|adam ok|grouped|tangent" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -14
done
