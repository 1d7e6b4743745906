# One box with N GPUs: the default bench line (as the driver runs it) [+ extra variants when $2 is set]
N=${1:-8}; TAG=${3:-r02g}
mkdir -p gpurun_out
if [ $N -eq 1 ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
(time timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench_n$N.json 2> gpurun_out/${TAG}_bench_n$N.err); tail -n 4 gpurun_out/${TAG}_bench_n$N.err
python tools/bench_summary.py gpurun_out/${TAG}_bench_n$N.json
if [ -n "$2" ]; then
  timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --only tiles --tiles-view-refit 0 > gpurun_out/${TAG}_tiles_refit0_n$N.json 2> gpurun_out/${TAG}_tiles_refit0_n$N.err
  python -c "
import json; t=json.load(open('gpurun_out/${TAG}_tiles_refit0_n$N.json'))['tiles']; print('tiles refit 0: ray', round(t['raycast']['value']), round(1e3*t['raycast']['ms_per_frame'],1), 'us/frame')"
fi
