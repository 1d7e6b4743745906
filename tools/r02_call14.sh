for sub in 8 16; do for st in 1; do RENDERTOY_B200_SUB=$sub timeout 200 python bench.py --steps 3 --warmup 3 --only raster 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raster']; print('SUB', sys.argv[1], 'value', round(x['value']), 'frame us', round(1e3*x['roofline']['frame_ms'],1), 'alone', round(1e3*x['roofline']['frame_ms_alone'],1), 'e2e', round(x['e2e']['value']))" $sub; done; done
for sub in 4 8 16; do RENDERTOY_B200_SUB=$sub timeout 200 python bench.py --steps 3 --warmup 3 --only raycast --raycast-streams $sub 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raycast']; print('raycast SUB=streams', sys.argv[1], 'value', round(x['value']), 'kernel us', round(1e3*x['roofline']['kernel_ms'],1), 'e2e', round(x['e2e']['value']))" $sub; done
