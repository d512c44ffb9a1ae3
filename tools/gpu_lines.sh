#!/bin/bash
# One-GPU bench lines of every workload with the current kernels (one gpurun call).  usage: tools/gpu_lines.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_reference_arm.json 2> $OUT/${TAG}_reference_arm.err
timeout 600 python bench.py --steps 10 --warmup 3 --api > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench.err
for w in cfg3 cfg5 active32; do
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 > $OUT/${TAG}_${w}_1gpu.json 2> $OUT/${TAG}_${w}.err
done
timeout 900 python bench.py --workload cfg4 --snowpacks 2500 --steps 1 --warmup 1 > $OUT/${TAG}_cfg4_1gpu.json 2> $OUT/${TAG}_cfg4.err
python - <<PY
import json
for w in ("bench", "cfg3", "cfg4", "cfg5", "active32"):
    try:
        d = json.load(open("$OUT/${TAG}_%s_1gpu.json" % w))
        c = d.get("cpu_baseline") or {}
        print(w, "value %.0f" % d["value"], "e2e %.0f" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"], d["roofline"]["kernel"],
              d["roofline"]["avg_launch_ms"], "cpu %.1f x%s err %s n %s" % (c.get("value", 0), c.get("cores"), c.get("max_rel_err"), c.get("n_compared")),
              "errors", d["errors"], d.get("e2e_api", {}).get("value"))
    except Exception as e:
        print(w, "no line", e)
PY
compute-sanitizer --tool racecheck --print-limit 50 python tools/sanitize_cases.py passive128 2>&1 | grep -E "SUMMARY|Error|finite" | head
