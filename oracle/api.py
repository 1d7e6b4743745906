"""oracle/api.py -- ctypes/numpy front end of liboracle.so (TEST INFRASTRUCTURE, NOT PRODUCT CODE)."""
import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

SHADER_LESSON08 = 8
SHADER_LESSON09 = 9
NO_WINNER = 0xFFFFFFFF


class _Config(C.Structure):
    _fields_ = [("shader", C.c_int), ("width", C.c_int), ("height", C.c_int),
                ("vs_globals", C.POINTER(C.c_float)), ("tex", C.POINTER(C.c_float)),
                ("tex_w", C.c_int), ("tex_h", C.c_int)]


class _Stats(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("triangles_in", "primitives", "skipped_z0", "dropped_large", "fragments",
                                        "fragments_offscreen", "tie_pixels", "pixels_written", "over_capacity")]


def build(force=False):
    """Compile liboracle.so with the committed Makefile (gcc, strict IEEE flags)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "raycast_oracle.c", "Makefile")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_bvh_build.restype = C.c_void_p
        _LIB.orc_num_threads.restype = C.c_int
    return _LIB


def num_threads():
    return int(lib().orc_num_threads())


def set_threads(n):
    """OpenMP threads for every following oracle call (bench.py: os.cpu_count(), whatever OMP_NUM_THREADS says)."""
    lib().orc_set_threads(int(n))
    return num_threads()


def _fp(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def _mesh_rows(mesh_vertices):
    m = np.ascontiguousarray(mesh_vertices)
    if m.dtype != np.float32:
        m = m.view(np.float32)
    return np.ascontiguousarray(m.reshape(-1, 20))


@dataclass
class RasterResult:
    depth: np.ndarray    # (H, W) uint32 -- bits of the float depth, like Raster.get_depth_buffer()
    bgra: np.ndarray     # (H, W, 4) uint8 -- CL_BGRA / UNORM_INT8 render target bytes
    winner: np.ndarray   # (H, W) uint32 -- primitive id 2*t+k owning the pixel after this draw, NO_WINNER if none
    tie: np.ndarray      # (H, W) uint8 -- 1 where more than one primitive produced the winning depth bits
    stats: dict


def draw_triangles(shader, width, height, mesh_vertices, vs_globals, indices=None, texture=None,
                   depth=None, bgra=None):
    """One Raster.draw_triangles (rendering/_raster.py:416-437) on cleared (or supplied) targets.
    mesh_vertices: (n, 20) float32 MeshVertex rows; vs_globals: 48 float32 (World, View, Proj)."""
    L = lib()
    mesh = _mesh_rows(mesh_vertices)
    g = np.ascontiguousarray(np.asarray(vs_globals, dtype=np.float32).reshape(48))
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    n_tris = (mesh.shape[0] if idx is None else idx.shape[0]) // 3
    if depth is None:
        depth = np.full((height, width), 0x3F800000, dtype=np.uint32)  # clear(depth, 1.0), lesson08:102
    if bgra is None:
        bgra = np.zeros((height, width, 4), dtype=np.uint8)             # clear(render_target), lesson08:101
    depth = np.ascontiguousarray(depth, dtype=np.uint32).copy()
    bgra = np.ascontiguousarray(bgra, dtype=np.uint8).copy()
    winner = np.full((height, width), NO_WINNER, dtype=np.uint32)
    tie = np.zeros((height, width), dtype=np.uint8)
    cfg = _Config(shader, width, height, _fp(g), None, 0, 0)
    tex = None
    if texture is not None:
        tex = np.ascontiguousarray(texture, dtype=np.float32)
        assert tex.ndim == 3 and tex.shape[2] == 4
        cfg.tex, cfg.tex_h, cfg.tex_w = _fp(tex), tex.shape[0], tex.shape[1]
    elif shader == SHADER_LESSON09:
        raise ValueError("lesson09 shader needs a texture")
    st = _Stats()
    rc = L.orc_draw_triangles(C.byref(cfg), _fp(mesh), _p(idx, C.c_int32), C.c_int64(n_tris), C.c_int64(mesh.shape[0]),
                              _p(depth, C.c_uint32), _p(bgra, C.c_uint8), _p(winner, C.c_uint32), _p(tie, C.c_uint8),
                              C.byref(st))
    assert rc == 0
    return RasterResult(depth, bgra, winner, tie, {n: int(getattr(st, n)) for n, _ in _Stats._fields_})


def draw_points(shader, width, height, mesh_vertices, vs_globals, indices=None, texture=None, depth=None, bgra=None):
    """One Raster.draw_points (rendering/_raster.py:399-414) on cleared (or supplied) targets."""
    L = lib()
    mesh = _mesh_rows(mesh_vertices)
    g = np.ascontiguousarray(np.asarray(vs_globals, dtype=np.float32).reshape(48))
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    n_points = mesh.shape[0] if idx is None else idx.shape[0]
    depth = np.full((height, width), 0x3F800000, dtype=np.uint32) if depth is None else np.ascontiguousarray(depth, dtype=np.uint32).copy()
    bgra = np.zeros((height, width, 4), dtype=np.uint8) if bgra is None else np.ascontiguousarray(bgra, dtype=np.uint8).copy()
    winner = np.full((height, width), NO_WINNER, dtype=np.uint32)
    cfg = _Config(shader, width, height, _fp(g), None, 0, 0)
    tex = None
    if texture is not None:
        tex = np.ascontiguousarray(texture, dtype=np.float32)
        cfg.tex, cfg.tex_h, cfg.tex_w = _fp(tex), tex.shape[0], tex.shape[1]
    st = _Stats()
    rc = L.orc_draw_points(C.byref(cfg), _fp(mesh), _p(idx, C.c_int32), C.c_int64(n_points), C.c_int64(mesh.shape[0]),
                           _p(depth, C.c_uint32), _p(bgra, C.c_uint8), _p(winner, C.c_uint32), C.byref(st))
    assert rc == 0
    return RasterResult(depth, bgra, winner, np.zeros((height, width), np.uint8), {n: int(getattr(st, n)) for n, _ in _Stats._fields_})


def vertex_kat(P, vs_globals, width, height):
    g = np.ascontiguousarray(np.asarray(vs_globals, dtype=np.float32).reshape(48))
    p = np.ascontiguousarray(P, dtype=np.float32)
    clip, scr = np.zeros(4, np.float32), np.zeros(4, np.float32)
    lib().orc_vertex_kat(_fp(p), _fp(g), width, height, _fp(clip), _fp(scr))
    return clip, scr


def _rays(rays):
    r = np.ascontiguousarray(rays, dtype=np.float32).reshape(-1, 8)
    return r


def _hit_arrays(n):
    return (np.empty(n, np.float32), np.empty(n, np.uint32), np.empty(n, np.float32), np.empty(n, np.float32))


def raycast_brute(mesh_vertices, rays, indices=None):
    mesh, r = _mesh_rows(mesh_vertices), _rays(rays)
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    n_tris = (mesh.shape[0] if idx is None else idx.shape[0]) // 3
    t, i, u, v = _hit_arrays(r.shape[0])
    lib().orc_raycast_brute(_fp(mesh), _p(idx, C.c_int32), C.c_int64(n_tris), _fp(r), C.c_int64(r.shape[0]),
                            _fp(t), _p(i, C.c_uint32), _fp(u), _fp(v))
    return t, i, u, v


def bvh_build(mesh_vertices, indices=None):
    mesh = _mesh_rows(mesh_vertices)
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    n_tris = (mesh.shape[0] if idx is None else idx.shape[0]) // 3
    return C.c_void_p(lib().orc_bvh_build(_fp(mesh), _p(idx, C.c_int32), C.c_int64(n_tris)))


def bvh_raycast(handle, rays):
    r = _rays(rays)
    t, i, u, v = _hit_arrays(r.shape[0])
    lib().orc_bvh_raycast(handle, _fp(r), C.c_int64(r.shape[0]), _fp(t), _p(i, C.c_uint32), _fp(u), _fp(v))
    return t, i, u, v


def bvh_free(handle):
    lib().orc_bvh_free(handle)


def primary_rays(cam, width, height, rect=None):
    x0, y0, w, h = rect if rect is not None else (0, 0, width, height)
    cam = np.ascontiguousarray(cam, dtype=np.float32).reshape(12)
    rays = np.empty((h * w, 8), np.float32)
    lib().orc_primary_rays(_fp(cam), width, height, x0, y0, w, h, _fp(rays))
    return rays


def shade_hits(mode, mesh_vertices, ids, u, v, indices=None, texture=None):
    mesh = _mesh_rows(mesh_vertices)
    idx = None if indices is None else np.ascontiguousarray(indices, dtype=np.int32)
    ids = np.ascontiguousarray(ids, np.uint32); u = np.ascontiguousarray(u, np.float32); v = np.ascontiguousarray(v, np.float32)
    out = np.empty((ids.shape[0], 4), np.uint8)
    tex, tw, th = None, 0, 0
    if texture is not None:
        tex = np.ascontiguousarray(texture, dtype=np.float32); th, tw = tex.shape[0], tex.shape[1]
    lib().orc_shade_hits(mode, _fp(mesh), _p(idx, C.c_int32), _p(ids, C.c_uint32), _fp(u), _fp(v), C.c_int64(ids.shape[0]),
                         _p(tex, C.c_float), tw, th, _p(out, C.c_uint8))
    return out
