#!/bin/bash
# quick A/B of prebuilt library variants: cfg-2 kernel times only.  usage: tools/_quick_ab.sh <tag> <lib.so> ...
TAG=$1; shift
for lib in "$@"; do
  cp $lib smrt_b200/csrc/libsmrt_dort_b200.so
  echo "=== $(basename $lib .so)" | tee -a gpurun_out/${TAG}_ab.log
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value', round(d['value']), 'ms', r['avg_launch_ms'])" | tee -a gpurun_out/${TAG}_ab.log
done
