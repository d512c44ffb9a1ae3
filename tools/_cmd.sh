mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > /dev/null 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --workload cfg5 --steps 2 --warmup 1 | tee gpurun_out/r01w_cfg5.json
