#!/bin/bash
# Multi-GPU bench lines of every BASELINE config (one gpurun --gpus N call).  usage: tools/gpu_multi.sh <tag> <N>
TAG=${1:-rXX}; N=${2:-8}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1
run() {  # name, extra args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N "$@" > $OUT/${TAG}_${name}_${N}gpu.json 2> $OUT/${TAG}_${name}_${N}gpu.err
  echo "== $name rc=$?"; python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_${name}_${N}gpu.json"))
    print("$name", "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "ms/step %.1f" % d["ms_per_step"], "errors", d["errors"], d["scaling"])
except Exception as e:
    print("$name: no line", e)
PY
}
run cfg2_weak --steps 5 --warmup 3 --no-cpu-baseline
run cfg2_strong --scaling strong --snowpacks 10000 --steps 5 --warmup 3 --no-cpu-baseline
run cfg5_full --workload cfg5 --steps 3 --warmup 3 --no-cpu-baseline
run cfg3 --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline
run cfg4 --workload cfg4 --snowpacks 2500 --steps 1 --warmup 1 --no-cpu-baseline
