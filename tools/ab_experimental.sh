#!/bin/bash
# Run on the GPU box (under gpurun, one GPU): first hardware run of the experimental traversal variants.
#   1. the parity test that is skipped by default (hits bit-identical with refit passes / region traversal on and off)
#   2. frame times of the cfg4 / lesson08 4K frames: measured path, refit x1/2/4/8, region traversal at three thresholds, both
#   3. one bench line with the best-looking combination (edit the flags), for the JSON record
# Output -> gpurun_out/ab_experimental.log
set -x
mkdir -p gpurun_out
{
RENDERTOY_B200_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_raycast_gpu.py -q -k experimental -s 2>&1 | tail -40
for refit in 0 1 2 4 8; do RT_VIEW_REFIT=$refit timeout 120 python tools/quick_raycast_bench.py ab 2>&1 | grep -E "refit|render"; done
for amax in 2 8 32; do
  RT_REGION_AMAX=$amax timeout 120 python tools/quick_raycast_bench.py ab 2>&1 | grep -E "region|render"
  RT_VIEW_REFIT=4 RT_REGION_AMAX=$amax timeout 120 python tools/quick_raycast_bench.py ab 2>&1 | grep -E "refit|region|render"
done
timeout 200 python bench.py --no-cpu-baseline --view-refit 4 --region-amax 8 2>/dev/null | cut -c1-400
} > gpurun_out/ab_experimental.log 2>&1
tail -60 gpurun_out/ab_experimental.log
