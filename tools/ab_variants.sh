#!/bin/bash
# Run on the GPU box: time the section given ($1 = raster|raycast|tiles) of bench.py with the default library and every variants/$2*.so
SEC=${1:-raster}; PAT=${2:-}
run() { RENDERTOY_B200_LIB=$1 timeout 300 python bench.py --only $SEC --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())[sys.argv[1]]
r=d['roofline']
print('%-28s value %9.0f  per-frame %.2f us  alone %.2f us  e2e %.0f' % (sys.argv[2], d['value'], 1e3*r.get('frame_ms', r.get('kernel_ms',0)), 1e3*r.get('frame_ms_alone', r.get('kernel_ms_alone',0)), d['e2e']['value']))" $SEC $2; }
run "" default
for lib in variants/$PAT*.so; do run $PWD/$lib $(basename $lib .so); done
run "" default-again
