"""oracle/ -- CPU checker for rendertoy_b200 (TEST INFRASTRUCTURE, NOT PRODUCT CODE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this package.  Nothing under rendertoy_b200/ does.
"""
from .api import (build, lib, draw_triangles, draw_points, RasterResult, raycast_brute, bvh_build, bvh_raycast, bvh_free,
                  primary_rays, shade_hits, vertex_kat, num_threads, set_threads, SHADER_LESSON08, SHADER_LESSON09, NO_WINNER)
from . import host_math
