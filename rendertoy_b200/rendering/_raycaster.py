"""Raycaster: closest-hit ray casting over triangle meshes (reference: rendering/_raycaster.py).

The reference declares the API and nothing else: BVH_AABB / BVH_Triangle structs (:8-22), a constructor whose
_build_ads() computes a triangle count and drops it (:30-33) and `ray_cast(rays) -> pass` (:35-36).  The names,
the struct fields and the (models) / (rays) signatures are kept; the behaviour is new and is defined by
oracle/raycast_oracle.c (float32 Moller-Trumbore, closest hit by (t bits, triangle id), two-sided):

  _build_ads()   GPU LBVH over all triangles of all models (rt_bvh_build: Morton codes -> radix sort ->
                 Karras hierarchy -> refit), mesh data replicated per GPU
  ray_cast(rays) rays: buffer of Ray {origin: float3, direction: float3}; returns a buffer of
                 RayHit {t, index, mesh, u, v} (miss: t = +inf, index = mesh = -1)
  render(...)    the fused hot path: primary rays of a camera for a pixel rect -> hits and/or shaded BGRA8
                 written straight into a render target (Lambert of lesson08:42 or texture of lesson09:90-95)
"""
import ctypes
import typing

import numpy as np
import torch

from .. import _native
from . import _core
from ._core import kernel_struct, float3, create_buffer, stream_ptr, DeviceBuffer
from ._modeling import Mesh

# render(): above this size the per-frame projection of the BVH (reads 64 B, writes 48 B per node) costs more than
# the cheaper node test saves
_CAM12 = ctypes.c_float * 12
_MAT16 = ctypes.c_float * 16
VIEW_NODES_MAX_TRIANGLES = 1 << 18
# default builder: clustering pays where the traversal is the bound (the screen-space packet path)
PLOC_MAX_TRIANGLES = 1 << 18
# render(): the scene's screen-space bound is the union of the projected bounds of this many chunks of the sorted leaves
SCREEN_CHUNKS = 64


@kernel_struct
class BVH_AABB:
    min: float3
    max: float3
    count: int
    elements: int


@kernel_struct
class BVH_Triangle:
    v0: float3
    v1: float3
    v2: float3
    index: int
    mesh: int


@kernel_struct
class Ray:
    origin: float3
    direction: float3


@kernel_struct
class RayHit:
    t: np.float32
    index: np.int32
    mesh: np.int32
    u: np.float32
    v: np.float32


def _mat16(x):
    """float4x4 value / (4, 4) array / 16 numbers -> (c_float * 16), row-major."""
    x = np.asarray(x)
    if x.dtype != _core.float4x4:
        x = np.ascontiguousarray(x, dtype=np.float32)
    assert x.nbytes == 64, "expected a 4x4 matrix"
    a = _MAT16()
    ctypes.memmove(a, x.ctypes.data, 64)
    return a


def camera_frame(view, proj, world=None):
    """{origin, U, V, W} (12 float32, model space) of the reference camera convention (rendering/_core.py:
    528-548): a pixel centre with NDC coordinates (sx, sy) looks along U*sx + V*sy + W.
    view/proj/world: float4x4 values (or 4x4 arrays), row-vector convention.  Computed in float64, rounded once
    (rt_camera_frame, host arithmetic in the native library: this runs once per frame and the same thing through numpy --
    two LAPACK inversions of tiny matrices -- costs more host time than the GPU needs for the 4K frame)."""
    out = np.empty(12, np.float32)
    ok = _native.lib().rt_camera_frame(_mat16(view), _mat16(proj), None if world is None else _mat16(world),
                                       out.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    if not ok:
        raise np.linalg.LinAlgError("camera_frame: singular view rotation or world matrix (or non-finite entries)")
    return out


class Raycaster:
    def __init__(self, models: typing.List[Mesh], builder: str = None):
        """builder: "ploc" (default up to PLOC_MAX_TRIANGLES: parallel locally-ordered clustering, the better tree) or
        "lbvh" (Karras hierarchy, the faster build).  Both give the same hits; the reference takes only `models`."""
        self.models = models
        self.builder = builder
        self._build_ads()

    def _build_ads(self):
        pos, nrm, idx, counts = [], [], [], []
        vbase = 0
        any_indexed = False
        for m in self.models:
            m: Mesh
            p4, n4 = _core.mesh_soa(m.vertices)
            nverts = m.vertices.shape[0]
            # only int32 index buffers are real (the kernels' [np.int32], _raster.py:154); load_obj's never-filled
            # `int` buffer (_loaders.py:17) is treated like the tutorials treat it: ignored
            use_idx = m.indices is not None and m.indices.dtype == np.int32
            triangles = nverts // 3 if not use_idx else m.indices.shape[0] // 3
            if use_idx:
                any_indexed = True
                i32 = m.indices.tensor().view(torch.int32)[:triangles * 3]
            else:
                i32 = None
            pos.append(p4[:nverts]); nrm.append(n4[:nverts]); idx.append((i32, vbase, triangles)); counts.append(triangles)
            vbase += nverts
        self.tri_offsets = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        self.n_triangles = int(self.tri_offsets[-1])
        assert self.n_triangles >= 1, "Raycaster needs at least one triangle"
        if len(self.models) == 1:
            self.pos4, self.nrm4 = pos[0], nrm[0]
            self.indices = idx[0][0]
        else:
            self.pos4, self.nrm4 = torch.cat(pos), torch.cat(nrm)
            if any_indexed:
                parts = []
                for i32, base, tris in idx:
                    parts.append((i32 if i32 is not None else torch.arange(tris * 3, dtype=torch.int32, device=self.pos4.device)) + base)
                self.indices = torch.cat(parts).to(torch.int32)
            else:
                self.indices = None
        # scene bounds on the host (one sync at build time): used for the screen-space cull rect and the FMA-slab guard
        xyz = self.pos4[:, :3]
        self.scene_lo = xyz.amin(0).double().cpu().numpy()
        self.scene_hi = xyz.amax(0).double().cpu().numpy()
        self.scene_extent = float((self.scene_hi - self.scene_lo).max())
        self._lo3, self._hi3 = (ctypes.c_double * 3)(*self.scene_lo.tolist()), (ctypes.c_double * 3)(*self.scene_hi.tolist())
        dev = self.pos4.device
        L = _native.lib()
        n = self.n_triangles
        self.nodes = torch.empty(int(L.rt_bvh_node_bytes(n)), dtype=torch.uint8, device=dev)
        self.tris = torch.empty(int(L.rt_bvh_tri_bytes(n)), dtype=torch.uint8, device=dev)
        scratch = torch.empty(int(L.rt_bvh_scratch_bytes(n)), dtype=torch.uint8, device=dev)
        if self.builder is None:
            self.builder = "ploc" if n <= PLOC_MAX_TRIANGLES else "lbvh"
        assert self.builder in ("ploc", "lbvh"), "builder must be 'ploc' or 'lbvh'"
        _native.call("rt_bvh_build", self.pos4.data_ptr(), self._idx_ptr(), n, self.nodes.data_ptr(), self.tris.data_ptr(),
                     scratch.data_ptr(), _native.BVH_PLOC if self.builder == "ploc" else _native.BVH_LBVH, stream_ptr())
        self._build_scratch = scratch  # kept until the stream has consumed it
        # Bounds of SCREEN_CHUNKS chunks of the leaves in builder (Morton) order, on the host (one more sync at build time): the
        # union of their projected rectangles is what render() culls with and returns as the frame's content rect -- a quarter
        # smaller than the rectangle of the scene's one box for the dragon orbit (0.228 -> 0.175 of the 4K frame).
        leaves = self.tris.view(torch.float32).view(-1, 12)[:n]
        v0, e1, e2 = leaves[:, 0:3], leaves[:, 4:7], leaves[:, 8:11]
        tri_lo = torch.minimum(v0, torch.minimum(v0 + e1, v0 + e2))
        tri_hi = torch.maximum(v0, torch.maximum(v0 + e1, v0 + e2))
        k = min(SCREEN_CHUNKS, n)
        m = -(-n // k)
        pad = k * m - n
        if pad:
            tri_lo = torch.cat([tri_lo, tri_lo[-1:].expand(pad, 3)])
            tri_hi = torch.cat([tri_hi, tri_hi[-1:].expand(pad, 3)])
        clo = tri_lo.view(k, m, 3).amin(1).double().cpu().numpy()
        chi = tri_hi.view(k, m, 3).amax(1).double().cpu().numpy()
        if not (np.isfinite(clo).all() and np.isfinite(chi).all()):      # non-finite data: the one scene box (no bound either, then)
            clo, chi = self.scene_lo[None], self.scene_hi[None]
        self._n_chunks = int(clo.shape[0])
        self._chunk_lo = (ctypes.c_double * (3 * self._n_chunks))(*clo.reshape(-1).tolist())
        self._chunk_hi = (ctypes.c_double * (3 * self._n_chunks))(*chi.reshape(-1).tolist())
        self._view_nodes = {}          # stream -> per-frame screen-space nodes of render() (scratch, allocated on first use)

    def _idx_ptr(self):
        return None if self.indices is None else self.indices.data_ptr()

    # -- generic rays ---------------------------------------------------------------------------------
    def ray_cast_native(self, rays: DeviceBuffer) -> torch.Tensor:
        """(n, 4) float32 tensor {t, triangle id bits, u, v} with GLOBAL triangle ids (no mesh split)."""
        n = rays.size
        assert rays.dtype.itemsize == 32, "rays must be Ray {origin: float3, direction: float3} (32 bytes)"
        hits = torch.empty((max(n, 1), 4), dtype=torch.float32, device=self.pos4.device)
        _native.call("rt_raycast_rays", self.nodes.data_ptr(), self.tris.data_ptr(), self.n_triangles, rays.ptr, n,
                     hits.data_ptr(), stream_ptr())
        return hits[:n]

    def ray_cast(self, rays: DeviceBuffer) -> DeviceBuffer:
        native = self.ray_cast_native(rays)
        n = native.shape[0]
        out = create_buffer(n, RayHit)
        t = out.tensor().view(torch.float32).view(n, 5)
        gid = native[:, 1].contiguous().view(torch.int32).to(torch.int64)
        miss = gid < 0  # 0xFFFFFFFF
        offs = torch.as_tensor(self.tri_offsets, device=native.device)
        mesh = torch.searchsorted(offs, gid.clamp(min=0), right=True) - 1
        index = gid - offs[mesh]
        mesh = torch.where(miss, torch.full_like(mesh, -1), mesh)
        index = torch.where(miss, torch.full_like(index, -1), index)
        t[:, 0] = native[:, 0]
        t[:, 1] = index.to(torch.int32).view(torch.float32)
        t[:, 2] = mesh.to(torch.int32).view(torch.float32)
        t[:, 3] = native[:, 2]
        t[:, 4] = native[:, 3]
        out.device_written()
        return out

    def screen_bounds(self, camera, W, H):
        """Conservative inclusive pixel rect [x0, y0, x1, y1] containing every pixel whose primary ray can hit the
        scene's bounding box (projected corners +- 2 px), or None when the box reaches behind the eye.
        Host arithmetic only, but native (rt_raycast_screen_bounds): it runs once per frame, and eight corners through
        numpy cost ~50 us of a ~120 us frame."""
        cam = camera if isinstance(camera, _CAM12) else _native.float_array_from_bytes(np.ascontiguousarray(camera, np.float32).reshape(12).view(np.uint8), 12)
        rect = (ctypes.c_int * 4)()
        if not _native.lib().rt_raycast_screen_bounds_n(cam, self._chunk_lo, self._chunk_hi, self._n_chunks, W, H, rect):
            return None
        return rect[0], rect[1], rect[2], rect[3]

    # -- fused primary rays -----------------------------------------------------------------------------
    def render(self, render_target, camera, rect=None, shader=_native.SHADER_LESSON08, texture_descriptor=None,
               hits: torch.Tensor = None, frame_size=None, stats: torch.Tensor = None, cull=True, view_nodes=None,
               stripes=None):
        """Primary rays for `rect` = (x0, y0, w, h) of the frame (default: the whole render target), closest hit,
        shade, write BGRA8 into `render_target` at the rect's position.  camera: 12 floats from camera_frame().
        hits: optional (h*w, 4) float32 tensor to also receive {t, id, u, v}.  render_target may be None when only
        hits are wanted (then frame_size=(W, H) is required).  stats: optional int64[3] tensor; the instrumented kernel
        adds {node visits, triangle tests, rays} to it.  cull: skip tracing outside the scene's projected bounds.
        view_nodes: project the BVH into this camera's screen space first and traverse that (default: yes up to
        VIEW_NODES_MAX_TRIANGLES triangles, where the per-frame projection pass pays for itself).
        stripes: (rows, mod, rem) -- the image-space partition of one frame over `mod` GPUs (parallel.tile_rects): only
        the row stripes s = y // rows with s % mod == rem are traced and written, in ONE launch (one projection pass per
        frame and rank, not one per band); all other pixels of the rect are left untouched.
        Returns the inclusive frame-pixel rect (x0, y0, x1, y1) outside of which everything this call wrote is the clear
        colour / a miss (the cull rect clipped to `rect`; x1 < x0 when nothing can be hit) -- what a sparse gather has to move."""
        if render_target is not None:
            W, H = render_target.width, render_target.height
        else:
            W, H = frame_size
        x0, y0, w, h = rect if rect is not None else (0, 0, W, H)
        tex = 0
        if shader == _native.SHADER_LESSON09:
            desc = texture_descriptor.get() if hasattr(texture_descriptor, "get") else texture_descriptor
            tex = _core.__MEMORY_POOL__.texture_handle(int(desc["offset"]))
        bgra_ptr = None
        if render_target is not None:
            assert render_target.is_bgra8, "render target must be the BGRA8 presenter image"
            # a full-frame render overwrites every pixel: a deferred clear is dropped, otherwise executed first
            if (x0, y0, w, h) == (0, 0, W, H) and (stripes is None or stripes[1] == 1):
                render_target.take_pending_clear()
            bgra_ptr = render_target.ptr + 4 * (y0 * W + x0)
        cam_c = _native.float_array_from_bytes(np.ascontiguousarray(camera, np.float32).reshape(12).view(np.uint8), 12)
        rect_c = None
        content = (x0, y0, x0 + w - 1, y0 + h - 1)
        if cull:
            r = (ctypes.c_int * 4)()
            if _native.lib().rt_raycast_screen_bounds_n(cam_c, self._chunk_lo, self._chunk_hi, self._n_chunks, W, H, r):
                rect_c = r
                content = (max(x0, r[0]), max(y0, r[1]), min(x0 + w - 1, r[2]), min(y0 + h - 1, r[3]))
        fast_slab = int(max(abs(cam_c[0]), abs(cam_c[1]), abs(cam_c[2])) <= 16.0 * self.scene_extent)
        if view_nodes is None:
            view_nodes = self.n_triangles <= VIEW_NODES_MAX_TRIANGLES
        vn_ptr = None
        if view_nodes:
            stream = stream_ptr()   # the scratch is rewritten by every call: one per stream, so frames on different streams may overlap
            if stream not in self._view_nodes:
                self._view_nodes[stream] = torch.empty(int(_native.lib().rt_raycast_view_node_bytes(self.n_triangles)), dtype=torch.uint8,
                                                       device=self.pos4.device)
            vn_ptr = self._view_nodes[stream].data_ptr()
        _native.call("rt_raycast_primary", self.nodes.data_ptr(), self.tris.data_ptr(), self.n_triangles, self.pos4.data_ptr(),
                     self.nrm4.data_ptr(), self._idx_ptr(),
                     cam_c,
                     W, H, x0, y0, w, h, shader, tex, None if hits is None else hits.data_ptr(), bgra_ptr, W,
                     None if stats is None else stats.data_ptr(), rect_c, fast_slab, vn_ptr,
                     None if stripes is None else (ctypes.c_int * 3)(*[int(v) for v in stripes]), stream_ptr())
        if render_target is not None:
            render_target._buffer.device_written()
        return content
