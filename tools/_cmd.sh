python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r01r_build.log 2>&1
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01r_bench.json 2> gpurun_out/r01r_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/r01r_bench.json'))
print(d['value'], d['roofline']['avg_launch_ms'], d['errors'])
PY
timeout 600 python bench.py --workload cfg3 --steps 2 --warmup 1 | tee gpurun_out/r01r_cfg3.json
timeout 900 python bench.py --workload cfg4 --steps 2 --warmup 1 | tee gpurun_out/r01r_cfg4.json
