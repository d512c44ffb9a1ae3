python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r01k_build.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for t in 512 256; do
  SMRT_B200_BOUNDARY_THREADS=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01k_bench_t$t.json 2>gpurun_out/r01k_bench_t$t.err
  python - <<PY
import json
d=json.load(open('gpurun_out/r01k_bench_t$t.json'))
print($t, d['value'], d['roofline']['avg_launch_ms'], d['errors'])
PY
done
