for tris in 2000 20000 50000 100000 200000; do RENDERTOY_B200_TRIS=$tris timeout 200 python bench.py --steps 3 --warmup 3 --only raster 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raster']; print('TRIS', sys.argv[1], 'value', round(x['value']), 'frame us', round(1e3*x['roofline']['frame_ms'],1), 'alone', round(1e3*x['roofline']['frame_ms_alone'],1), 'e2e', round(x['e2e']['value']))" $tris; done
