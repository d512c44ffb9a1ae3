#!/bin/bash
# usage: tools/variant_session.sh <tag> <lib1.so> [<lib2.so> ...]: for each prebuilt library variant (built in the
# authoring container under smrt_b200/csrc/variants/), install it as the product library on the (scratch) GPU box,
# print fixture errors and a short bench line.
TAG=$1; shift
OUT=gpurun_out; mkdir -p $OUT
for lib in "$@"; do
  name=$(basename $lib .so)
  cp $lib smrt_b200/csrc/libsmrt_dort_b200.so
  echo "=== $name" | tee -a $OUT/${TAG}_variants.log
  timeout 300 python tools/fixture_errors.py 2>&1 | tail -3 | tee -a $OUT/${TAG}_variants.log
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms', r['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])" | tee -a $OUT/${TAG}_variants.log
  if [[ -n "$VARIANT_CFG4" ]]; then
    timeout 300 python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('cfg4 value', round(d['value']), 'ms', r['avg_launch_ms'], 'clk', d['clocks']['sm_mhz'])" | tee -a $OUT/${TAG}_variants.log
  fi
done
