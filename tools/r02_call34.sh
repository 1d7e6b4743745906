timeout 120 python -m pytest tests/test_raster_gpu.py -m gpu -x -q -k "pinned_host or tile_push" 2>&1 | tail -3
for rb in tiles rect; do
timeout 120 python bench.py --steps 3 --warmup 3 --only raycast --readback $rb 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raycast']; e=x['e2e']; print('raycast readback', sys.argv[1], 'e2e', round(e['value']), 'd2h MB/frame', round(e['d2h_bytes_per_step']/512/1e6,2), 'value', round(x['value']))" $rb
timeout 120 python bench.py --steps 3 --warmup 3 --only raster --readback $rb 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raster']; e=x['e2e']; print('raster readback', sys.argv[1], 'e2e', round(e['value']), 'd2h MB/frame', round(e['d2h_bytes_per_step']/1024/1e6,2), 'value', round(x['value']))" $rb
done
