#!/bin/bash
# End-of-round session in ONE gpurun call: GPU tests, smoke, bench lines of every workload (tools/gpu_lines.sh), then the
# ncu launch lists and --set full captures of the cfg-2 and cfg-4 kernels reduced on the box to the CSV pages the
# summaries need, and the sanitizer passes over the cases that run the 64 < h <= 128 kernels.  usage: tools/gpu_final.sh <tag>
TAG=${1:-rXX}
OUT=gpurun_out; mkdir -p $OUT
bash tools/gpu_lines.sh $TAG
cap() {  # name, bench args
  local name=$1; shift
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'eigen_kernel|boundary_kernel' -c 2 -f \
    -o /tmp/${TAG}_${name} python bench.py "$@" --steps 1 --warmup 0 --no-cpu-baseline > $OUT/${TAG}_${name}_ncu_full.log 2>&1
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page raw --csv > $OUT/${TAG}_${name}_raw.csv 2>/dev/null
  ncu -i /tmp/${TAG}_${name}.ncu-rep --page source --csv --print-source cuda,sass > /tmp/${TAG}_${name}_cs.csv 2>/dev/null
  python tools/ncu_hotspots.py /tmp/${TAG}_${name}_cs.csv 30 > $OUT/${TAG}_${name}_hotspots.txt 2>&1
  for k in eigen boundary; do
    echo "##### top SASS instructions by stall samples, $k kernel" >> $OUT/${TAG}_${name}_hotspots.txt
    python tools/ncu_sass_top.py /tmp/${TAG}_${name}_cs.csv $k 20 >> $OUT/${TAG}_${name}_hotspots.txt 2>&1
  done
}
cap cfg2 --snowpacks 256
cap cfg4 --workload cfg4 --snowpacks 13
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_cfg2_launches.csv \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_cfg2_ncu_launch.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_cfg4_launches.csv \
  python bench.py --workload cfg4 --snowpacks 150 --steps 1 --warmup 1 --no-cpu-baseline > $OUT/${TAG}_cfg4_ncu_launch.log 2>&1
bash tools/gpu_sanitize.sh $TAG active32 passive64
du -sh $OUT
