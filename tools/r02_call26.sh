N=${1:-1}
if [ $N -eq 1 ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
for g in 1 0; do
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --only tiles --tiles-graphs $g > gpurun_out/r02j_tiles_g${g}_n$N.json 2> gpurun_out/r02j_tiles_g${g}_n$N.err
python -c "
import json; t=json.load(open('gpurun_out/r02j_tiles_g${g}_n$N.json'))['tiles']
for k in ('raycast','raster'): print('N=$N graphs=$g', k, round(t[k]['value']), round(1e3*t[k]['ms_per_frame'],1), 'us/frame', t[k]['gathered_frame_verified'], '|', t[k]['launch'][:110])" || tail -n 15 gpurun_out/r02j_tiles_g${g}_n$N.err
done
