# Run on the GPU box: ncu --set full over every kernel ONCE (profiler ranges; torch's own kernels excluded), summaries into profiles/
TAG=${1:-r02}
mkdir -p gpurun_out/profiles
NOT_TORCH='regex:^(?!vectorized_|elementwise_|unrolled_|reduce_|index|Cat|cat|triu|distribution).*$'
(time timeout 700 ncu --profile-from-start off --set full --clock-control none --import-source on -k "$NOT_TORCH" -c 80 -f -o /tmp/${TAG}_all python tools/profile_all.py) > gpurun_out/${TAG}_profile_all.log 2>&1
grep -c "PROF== Profiling" gpurun_out/${TAG}_profile_all.log; tail -4 gpurun_out/${TAG}_profile_all.log
ncu -i /tmp/${TAG}_all.ncu-rep --page raw --csv > gpurun_out/profiles/${TAG}_raw.csv 2>/dev/null
python tools/profile_inventory.py /tmp/${TAG}_all.ncu-rep $TAG > gpurun_out/${TAG}_inventory.log 2>&1
tail -45 gpurun_out/${TAG}_inventory.log
for k in raster_kernel coverage_kernel resolve_kernel raycast_kernel project_kernel view_refit_kernel fill_u64_kernel; do
  timeout 120 python tools/ncu_lines.py /tmp/${TAG}_all.ncu-rep $k 16 > gpurun_out/profiles/${TAG}_lines_$k.txt 2>&1
done
(time timeout 300 ncu --set full --clock-control none -k 'regex:karras|leaves_refit|ploc_leaves|ploc_flag|ploc_scan|bounds_init' -c 9 -f -o /tmp/${TAG}bvh python tools/profile_all.py bvh) > gpurun_out/${TAG}_profile_bvh.log 2>&1
python tools/profile_inventory.py /tmp/${TAG}bvh.ncu-rep ${TAG}bvh > gpurun_out/${TAG}_inventory_bvh.log 2>&1
tail -20 gpurun_out/${TAG}_inventory_bvh.log
cp profiles/${TAG}* profiles/traffic.json gpurun_out/profiles/
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_bench_${TAG}.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-config4 --no-tiles > gpurun_out/launches_bench_${TAG}.log 2>&1
ls gpurun_out/profiles | wc -l
