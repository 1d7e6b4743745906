"""`rendering` -- the reference package's public API (rendering/__init__.py:1-15), served by rendertoy_b200.

A tutorial script keeps its `import rendering as ren` (the repo-root `rendering/` shim re-exports this
package) and every name below keeps the reference's signature.
"""
from ._core import kernel_main, kernel_struct, create_buffer_from, create_buffer, create_struct, mapped, \
    kernel_function, create_struct_from, create_image2d, Image, Buffer, \
    r_image1d_t, w_image1d_t, r_image2d_t, w_image2d_t, r_image3d_t, w_image3d_t, \
    make_float2, make_float3, make_float4, make_float4x4, translate, identity, scale, rotate, matmul, to_array, clear, \
    perspective, look_at, normalize, dot

from ._core import float2, float3, float4, int2, int3, int4, uint2, uint3, uint4, float4x4, Texture2D, create_texture2D

from ._modeling import Mesh, WeldMode, SubdivisionMode, MeshVertex, manifold

from ._loaders import load_obj

from ._presentation import create_presenter, Presenter, Event

from ._raster import Raster

from ._raycaster import Raycaster

from . import _core, _modeling, _loaders, _presentation, _raster, _raycaster, _dsl
