for mc in 8 32; do for sub in 8 16; do CUDA_DEVICE_MAX_CONNECTIONS=$mc RENDERTOY_B200_SUB=$sub timeout 200 python bench.py --steps 3 --warmup 3 --only raster 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raster']; print('MAX_CONNECTIONS', sys.argv[2], 'SUB', sys.argv[1], 'value', round(x['value']), 'frame us', round(1e3*x['roofline']['frame_ms'],1), 'alone', round(1e3*x['roofline']['frame_ms_alone'],1), 'e2e', round(x['e2e']['value']))" $sub $mc; done; done
CUDA_DEVICE_MAX_CONNECTIONS=32 timeout 200 python bench.py --steps 3 --warmup 3 --only raycast --raycast-streams 8 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raycast']; print('raycast MAX_CONNECTIONS 32 streams 8', 'value', round(x['value']), 'kernel us', round(1e3*x['roofline']['kernel_ms'],1), 'e2e', round(x['e2e']['value']))"
