timeout 300 python -m pytest tests/test_raycast_gpu.py -m gpu -x -q -k "stripe_push" 2>&1 | tail -3
timeout 200 python tools/host_overhead_raycast.py 2>&1 | grep -v "^$" | head -45
