"""rendertoy_b200 -- B200-native (sm_100a) implementation of RenderToy's two per-pixel hot paths
(Raster.draw_triangles and Raycaster.ray_cast) behind the reference's `rendering` package API.

    rendertoy_b200.rendering   host-side mirror of the reference package (import as `rendering`)
    rendertoy_b200.csrc        hand-written CUDA kernels + the C ABI (include/rendertoy_b200.h)
    rendertoy_b200._native     ctypes binding of librendertoy_b200.so (fails loudly if it is missing)
    rendertoy_b200.scenes      synthetic stand-ins for the reference's missing models/dragon.obj
"""
__version__ = "0.1.0"
