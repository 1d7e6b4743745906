# N=2 box: full bench, then the alternative gather modes
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
(time timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_bench_n$N.json 2> gpurun_out/r02d_bench_n$N.err); tail -3 gpurun_out/r02d_bench_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --only raster --raster-gather peer > gpurun_out/r02d_raster_peer_n$N.json 2> gpurun_out/r02d_raster_peer_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --only tiles --tiles-gather copy > gpurun_out/r02d_tiles_copy_n$N.json 2> gpurun_out/r02d_tiles_copy_n$N.err
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --only raycast --commit-every 1 > gpurun_out/r02d_ray_c1_n$N.json 2> gpurun_out/r02d_ray_c1_n$N.err
python - $N <<'PY'
import json,sys
N=sys.argv[1]
def load(p):
    try: return json.load(open(p))
    except Exception as e: print(p, "FAILED", e); return None
d=load(f"gpurun_out/r02d_bench_n{N}.json")
if d:
    print("ray", round(d["value"]), "Mrays/s e2e", round(d["e2e"]["value"]), d["run"]["timed_region_ms_per_rank"])
    s=d["secondary"]; print("ras(copy)", round(s["value"]), "Mtris/s e2e", round(s["e2e"]["value"]), s["run"]["timed_region_ms_per_rank"])
    t=d["tiles"]; print("tiles(peer) ray", round(t["raycast"]["value"]), t["raycast"]["ms_per_frame"], "ras", round(t["raster"]["value"]), t["raster"]["ms_per_frame"])
    c=d["config4"]; print("cfg4 ras", round(c["raster"]["value"]), "ray", round(c["raycast"]["value"]))
x=load(f"gpurun_out/r02d_raster_peer_n{N}.json")
if x: print("ras(peer)", round(x["raster"]["value"]), x["raster"]["run"]["timed_region_ms_per_rank"])
x=load(f"gpurun_out/r02d_tiles_copy_n{N}.json")
if x: t=x["tiles"]; print("tiles(copy) ray", round(t["raycast"]["value"]), t["raycast"]["ms_per_frame"], "ras", round(t["raster"]["value"]), t["raster"]["ms_per_frame"])
x=load(f"gpurun_out/r02d_ray_c1_n{N}.json")
if x: print("ray commit-every 1", round(x["raycast"]["value"]), x["raycast"]["run"]["timed_region_ms_per_rank"])
PY
for f in gpurun_out/r02d_*_n$N.err; do tail -n 3 $f; done
