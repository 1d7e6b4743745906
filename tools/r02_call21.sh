N=8
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_raster_gpu.py -m gpu -x -q -k "tile_push" 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
for push in tiles rect; do
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --only raster --raster-push $push > gpurun_out/r02h_raster_${push}_n$N.json 2> gpurun_out/r02h_raster_${push}_n$N.err
python -c "
import json; x=json.load(open('gpurun_out/r02h_raster_${push}_n$N.json'))['raster']; print('raster push $push:', round(x['value']), 'Mtris/s', round(1e3*x['roofline']['frame_ms'],1), 'us/frame/rank; gather bytes/step/rank', [round(b/1e6) for b in x['run'].get('gather_bytes_per_step_per_rank',[])], 'of', round(x['run'].get('full_frame_bytes_per_step_per_rank',0)/1e6), 'MB; e2e', round(x['e2e']['value']))"
tail -n 2 gpurun_out/r02h_raster_${push}_n$N.err
done
