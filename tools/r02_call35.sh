timeout 200 python -m pytest tests/test_raycast_gpu.py tests/test_raster_gpu.py tests/test_edge_cases_gpu.py -m gpu -x -q -k "content_rect or sparse_push or stripe or scissor or tile or full_size or degenerate or edge" 2>&1 | tail -3
timeout 120 python bench.py --steps 3 --warmup 3 --only raycast 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raycast']; e=x['e2e']; print('raycast value', round(x['value']), 'us/frame', round(1e3*x['roofline']['kernel_ms'],1), 'traced', round(x['roofline']['fp32']['rays_traced_fraction'],3), 'e2e', round(e['value']), 'd2h MB/frame', round(e['d2h_bytes_per_step']/512/1e6,2), 'dense', round(e['dense_readback']['value']))"
timeout 120 python bench.py --steps 3 --warmup 3 --only raster 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raster']; e=x['e2e']; print('raster value', round(x['value']), 'e2e', round(e['value']), 'd2h MB/frame', round(e['d2h_bytes_per_step']/1024/1e6,2))"
