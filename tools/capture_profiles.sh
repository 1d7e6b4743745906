#!/bin/bash
# Run on the GPU box (under gpurun): launch lists + one full capture per hot kernel -> gpurun_out/
set -x
TAG=${1:-r01}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_raster_${TAG}.csv python tools/quick_raster_bench.py ncu > gpurun_out/launches_raster_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_raycast_${TAG}.csv python tools/quick_raycast_bench.py ncu > gpurun_out/launches_raycast_${TAG}.log 2>&1
for k in raster_kernel coverage_kernel resolve_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/prof_${k}_${TAG} python tools/quick_raster_bench.py ncu > /dev/null 2>&1
done
ncu --set full --clock-control none --import-source on -k regex:raycast_kernel -s 3 -c 1 -o gpurun_out/prof_raycast_kernel_${TAG} python tools/quick_raycast_bench.py ncu > /dev/null 2>&1
ls -la gpurun_out/
