mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_raycast_gpu.py -m gpu -x -q -k "refit or quoted" 2>&1 | tail -3
for r in 0 2 4 8 16 32; do timeout 200 python bench.py --steps 3 --warmup 3 --only raycast --view-refit $r 2>/dev/null | python -c "
import json,sys
x=json.loads(sys.stdin.read())['raycast']; print('refit iters', sys.argv[1], 'value', round(x['value']), 'kernel_ms', round(1e3*x['roofline']['kernel_ms'],1), 'alone', round(1e3*x['roofline']['kernel_ms_alone'],1), 'visits/ray', round(x['roofline']['fp32']['inner_node_visits_per_ray'],2), 'ff us', round(1e3*x['frame_filling']['ms_per_frame'],1), 'ff visits', round(x['frame_filling']['inner_node_visits_per_ray'],2))" $r; done
