"""Mesh container, MeshVertex layout and the `manifold` grid generator (reference: rendering/_modeling.py).

MeshVertex keeps the reference's 80-byte AoS layout (P@0 N@16 C@32 T@48 B@64, _modeling.py:22-28) because
tutorial code writes it through mapped(); Raster/Raycaster convert it once to SoA float4 arrays on the device
(rt_mesh_upload_soa) and cache that per buffer version.
"""
from enum import IntEnum

import numpy as np

from ._core import kernel_struct, float3, float2, create_buffer, mapped


class WeldMode(IntEnum):
    NONE = 0
    POSITION = 1
    POSITION_NORMAL_TEXTURE = 2
    ALL_ATTRIBUTES = 3


class SubdivisionMode(IntEnum):
    NONE = 0
    FLAT = 1
    LOOP = 2
    BUTTERFLY = 3


@kernel_struct
class MeshVertex:
    P: float3
    N: float3
    C: float2
    T: float3
    B: float3


class Mesh:
    """(vertices, indices) holder.  The editing operations are declared but unimplemented in the reference
    (_modeling.py:36-52) and stay so here: same exception, same message."""

    def __init__(self, vertices, indices):
        self.vertices = vertices
        self.indices = indices

    def _todo(self, *a, **k):
        raise Exception('Not implemented yet')

    clone = weld = simplify = subdivide = compute_normals = compute_tangents = _todo


def manifold(slices, stacks) -> 'Mesh':
    """(slices+1) x (stacks+1) unit grid in the z=0 plane, UV = xy, slices*stacks*2 indexed triangles
    (reference: _modeling.py:64-103).  The index rows advance by `slices`, not `slices+1`, exactly as the
    reference does (its own comments call this out); kept so meshes match."""
    u = np.arange(0, 1.0 + 0.5 / slices, 1.0 / slices)
    v = np.arange(0, 1.0 + 0.5 / stacks, 1.0 / stacks)
    vertices = create_buffer((slices + 1) * (stacks + 1), MeshVertex)
    indices = create_buffer(slices * stacks * 6, np.int32)
    floats_per_vertex = MeshVertex.itemsize // 4
    with mapped(vertices) as vmap:
        rows = vmap.ravel().view(np.float32).reshape(-1, floats_per_vertex)
        gu, gv = np.meshgrid(u, v)              # v is the slow axis, u the fast one
        rows[:, 0] = gu.ravel(); rows[:, 1] = gv.ravel(); rows[:, 2] = 0.0
        rows[:, 8] = rows[:, 0]; rows[:, 9] = rows[:, 1]
    with mapped(indices) as imap:
        col = np.arange(slices)[None, :]
        lo = col + np.arange(stacks)[:, None] * slices      # c00 of every cell, per stack row
        hi = lo + slices                                    # c10
        first = np.stack([lo, lo + 1, hi + 1], axis=-1).reshape(stacks, -1)    # (c00, c01, c11)
        second = np.stack([lo, hi + 1, hi], axis=-1).reshape(stacks, -1)       # (c00, c11, c10)
        imap[:] = np.concatenate([first, second], axis=1).ravel()
    return Mesh(vertices, indices)
