mkdir -p gpurun_out
ls -R oracle/_ref | head -20
timeout 300 python -m pytest tests/test_tutorials_gpu.py -m gpu -x -q -rs 2>&1 | tail -15
(time timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02c_bench_n1.json 2> gpurun_out/r02c_bench_n1.err); tail -5 gpurun_out/r02c_bench_n1.err
for r in 1 2 4; do timeout 300 python bench.py --steps 5 --warmup 3 --only raycast --view-refit $r > gpurun_out/r02c_refit$r.json 2>gpurun_out/r02c_refit$r.err; done
python - <<'PY'
import json
d=json.load(open("gpurun_out/r02c_bench_n1.json"))
print("ray", round(d["value"]), "Mrays/s  kernel_ms", d["roofline"]["kernel_ms"], "alone", d["roofline"]["kernel_ms_alone"], "e2e", round(d["e2e"]["value"]), "dense", round(d["e2e"]["dense_readback"]["value"]), "ff", d["frame_filling"]["value"], d["frame_filling"]["ms_per_frame"])
s=d["secondary"]; print("ras", round(s["value"]), "Mtris/s frame_ms", s["roofline"]["frame_ms"], "alone", s["roofline"]["frame_ms_alone"], "e2e", round(s["e2e"]["value"]))
t=d["tiles"]; print("tiles ray", round(t["raycast"]["value"]), t["raycast"]["ms_per_frame"], "ras", round(t["raster"]["value"]), t["raster"]["ms_per_frame"])
c=d["config4"]; print("cfg4 ras", round(c["raster"]["value"]), c["raster"]["ms_per_frame_per_rank"], c["raster"]["roofline"]["frac"], "ray", round(c["raycast"]["value"]), c["raycast"]["ms_per_frame_per_rank"], c["raycast"]["roofline"]["frac"], "gen", c["scene_generation_s"])
print("cpu", d.get("cpu_baseline"), s.get("cpu_baseline"))
print("clocks", d["clocks"], "ms_per_step", d["ms_per_step"])
for r in (1,2,4):
    try:
        x=json.load(open(f"gpurun_out/r02c_refit{r}.json"))["raycast"]; print("refit",r, round(x["value"]), x["roofline"]["kernel_ms"], x["roofline"]["kernel_ms_alone"], "ff", x["frame_filling"]["ms_per_frame"])
    except Exception as e: print("refit", r, "failed", e)
PY
