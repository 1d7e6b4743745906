N=${1:-2}; TAG=${2:-r02k}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR bench.py --gpus $N --steps 20 --warmup 5 --only tiles > gpurun_out/${TAG}_tiles_n$N.json 2> gpurun_out/${TAG}_tiles_n$N.err
python -c "
import json; t=json.load(open('gpurun_out/${TAG}_tiles_n$N.json'))['tiles']
for k in ('raycast','raster'): print('N=$N', k, round(t[k]['value']), round(1e3*t[k]['ms_per_frame'],1), 'us/frame', t[k]['gathered_frame_verified'], t[k]['streams'], t[k].get('view_refit_iterations'))" || tail -n 15 gpurun_out/${TAG}_tiles_n$N.err
