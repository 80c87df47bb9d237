#!/bin/bash
# compute-sanitizer evidence (run under gpurun): racecheck, memcheck and synccheck over tools/sanitize_cases.py.
# Output: gpurun_out/sanitizer/<tool>.log (copied to profiles/ by hand).
set -u
O=gpurun_out/sanitizer; mkdir -p $O
for tool in memcheck racecheck synccheck; do
  timeout ${SAN_TIMEOUT:-900} compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py "$@" > $O/$tool.log 2>&1
  echo "$tool rc=$?" >> $O/$tool.log
  tail -4 $O/$tool.log
done
