#!/bin/bash
# compute-sanitizer passes over tools/sanitize_cases.py; logs under gpurun_out/<tag>_sanitizer_<tool>.log
TAG=${1:-rXX}; shift
CASES="$*"
OUT=gpurun_out; mkdir -p $OUT
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 400 python tools/sanitize_cases.py $CASES > $OUT/${TAG}_sanitizer_${tool}.log 2>&1
  echo "== $tool rc=$?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|finite=|Error:" $OUT/${TAG}_sanitizer_${tool}.log | sort | uniq -c | sort -rn | head -30
done
