N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
i=0
for extra in "--raycast-streams 8" "--raycast-streams 8 --tiles-view-refit 2" "--raycast-streams 2"; do
i=$((i+1))
timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --only tiles $extra > gpurun_out/r02j_tiles_v${i}_n$N.json 2> gpurun_out/r02j_tiles_v${i}_n$N.err
python -c "
import json; t=json.load(open('gpurun_out/r02j_tiles_v${i}_n$N.json'))['tiles']
for k in ('raycast','raster'): print('N=$N $extra:', k, round(t[k]['value']), round(1e3*t[k]['ms_per_frame'],1), 'us/frame', t[k]['gathered_frame_verified'])" || tail -n 15 gpurun_out/r02j_tiles_v${i}_n$N.err
done
