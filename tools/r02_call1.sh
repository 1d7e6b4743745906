mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r02a_smi.txt
(time timeout 900 python -m pytest tests -m gpu -x -q) > gpurun_out/r02a_pytest.log 2>&1
bash tools/ab_experimental.sh > /dev/null 2>&1
timeout 200 python tools/quick_raster_bench.py > gpurun_out/r02a_quick_raster.log 2>&1
timeout 200 python tools/host_overhead.py > gpurun_out/r02a_host_overhead.log 2>&1
tail -5 gpurun_out/r02a_pytest.log; cat gpurun_out/ab_experimental.log | tail -60; cat gpurun_out/r02a_quick_raster.log; head -40 gpurun_out/r02a_host_overhead.log
