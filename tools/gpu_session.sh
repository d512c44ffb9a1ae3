#!/bin/bash
# One gpurun call: GPU parity tests, bench line, ncu launch list of the bench command, ncu --set full of both kernels.
# usage: tools/gpu_session.sh <tag> [flags: notests nolaunch nofull nobench]
TAG=${1:-rXX}
FLAGS=" $* "
OUT=gpurun_out
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" > $OUT/${TAG}_build.log 2>&1 || tail -20 $OUT/${TAG}_build.log
if [[ "$FLAGS" != *" notests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1
  tail -5 $OUT/${TAG}_pytest_gpu.log
fi
if [[ "$FLAGS" != *" nobench "* ]]; then
  timeout 600 python bench.py --steps 3 --warmup 3 > $OUT/${TAG}_bench_1gpu.json 2> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench_1gpu.json; tail -5 $OUT/${TAG}_bench.err
fi
if [[ "$FLAGS" != *" nolaunch "* ]]; then
  # launch list of the same command (cold-cache, serialised: shares only)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_ncu_launch_bench.log 2>&1
fi
if [[ "$FLAGS" != *" nofull "* ]]; then
  # full capture of one launch of each kernel on a reduced batch (256 snowpacks x 6 frequencies)
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'eigen_kernel|boundary_kernel' -c 2 \
    -f -o $OUT/${TAG}_prof python bench.py --snowpacks 256 --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
fi
ls -la $OUT | tail -12
