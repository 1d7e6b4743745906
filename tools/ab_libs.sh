#!/bin/bash
# Dev-time: time the 4K lesson06 frame with every variant library under variants/ (built by hand with -D knobs).
echo -n "default: "; python tools/quick_raycast_bench.py 2>&1 | head -1
for lib in variants/*.so; do echo -n "$lib: "; RENDERTOY_B200_LIB=$lib python tools/quick_raycast_bench.py 2>&1 | head -1; done
