mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_raster_gpu.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --steps 5 --warmup 3 --only raster > gpurun_out/r02e_raster_n1.json 2> gpurun_out/r02e_raster_n1.err; tail -3 gpurun_out/r02e_raster_n1.err
python - <<'PY'
import json
x=json.load(open("gpurun_out/r02e_raster_n1.json"))["raster"]; print("raster", round(x["value"]), x["roofline"]["frame_ms"], "alone", x["roofline"]["frame_ms_alone"], "e2e", round(x["e2e"]["value"]))
PY
timeout 200 python tools/host_overhead.py 2>&1 | head -12
