"""Raster: the reference's streaming rasterizer API (rendering/_raster.py:353-437) on native sm_100a kernels.

Same constructor, same methods, same assertion messages.  What is behind them differs:

  reference draw_triangles (:416-437)                         here
  -------------------------------------------------------    ---------------------------------------------
  VertexProcess -> TriangleAssembly -> (host sync) ->          ONE C-ABI call, rt_raster_draw_triangles:
  Dehomogenize -> loop{TriangleRaster -> (host sync) ->        fused vertex/clip/setup/coverage kernel with
  DepthTest -> FragmentProcess}; 32*W*H-fragment stream,       64-bit atomicMin on (depth<<32 | primitive),
  >= 3 blocking read-backs and >= 5 allocations per draw       then a resolve/shade kernel; no host sync,
                                                               no fragment stream, no per-draw allocation
  depth buffer: W*H uint32                                     high words of the W*H uint64 key buffer
  shaders: OpenCL C compiled at first dispatch                 tutorial shader pairs recognised by source
                                                               fingerprint -> hand-written CUDA twins

Results are bit-identical to the oracle's restatement of the reference pipeline (depth bits, winners, BGRA8).
"""
import hashlib
import os
import re
from enum import IntEnum

import numpy as np

from .. import _native
from . import _core
from ._core import float4, create_buffer, DepthView, stream_ptr


class FillMode(IntEnum):
    NONE = 0
    POINTS = 1
    WIREFRAME = 2
    SOLID = 3


def shader_fingerprint(source: str) -> str:
    """Whitespace- and comment-insensitive digest of a shader body."""
    s = re.sub(r"/\*.*?\*/", " ", source or "", flags=re.S)
    s = re.sub(r"//[^\n]*", " ", s)
    s = re.sub(r"\s+", "", s)
    return hashlib.sha256(s.encode()).hexdigest()[:16]


# (vertex fingerprint, fragment fingerprint) -> native shader id.  Values printed by tools/shader_fingerprints.py
# from tutorials/lesson08_rasterization.py:36-62 and tutorials/lesson09_texture_mapping.py:67-95.
_BUILTIN_SHADERS = {
    ("8952b904753dc90f", "b09f44910652fc57"): _native.SHADER_LESSON08,
    ("6fc2010a5ffe2329", "ec857604fe3a70ad"): _native.SHADER_LESSON09,
}


def transforms48(world, view, proj):
    """(World, View, Proj) -> 48 float32, row-major: each matrix may be a float4x4 value, the 16-tuple make_float4x4 /
    matmul return for array arguments (rendering/_core.py:127-181), or any 4x4 / 16-element array."""
    out = np.empty(48, np.float32)
    for i, m in enumerate((world, view, proj)):
        a = np.asarray(m)
        a = a.reshape(-1).view(np.float32) if a.dtype == _core.float4x4 else np.asarray(a, dtype=np.float32).reshape(-1)
        assert a.size == 16, "expected a 4x4 matrix"
        out[16 * i:16 * i + 16] = a
    return out


def _struct_fields(dtype):
    return [(n, dtype.fields[n][0], dtype.fields[n][1]) for n in dtype.names]


def _check_layout(dtype, expected, what):
    got = [(str(np.dtype(t)), off) for _, t, off in _struct_fields(dtype)]
    want = [(str(np.dtype(t)), off) for t, off in expected]
    if got != want:
        raise NotImplementedError(f"{what} layout {got} does not match the built-in shader's {want}")


class Raster:

    def __init__(self, render_target, vertex_shader, vertex_shader_globals, fragment_shader, fragment_shader_globals):
        # both resolve paths store packed BGRA8 words (the presenter's CL_BGRA / UNORM_INT8 image, _core.py:340)
        assert getattr(render_target, "is_bgra8", False), "Raster renders into a BGRA8 image (create_image2d(w, h, RGBA))"
        self._render_target = render_target
        n_pixels = render_target.width * render_target.height
        # depth lives in the high word of a 64-bit key per pixel; starts at 0 like the reference's zero-filled
        # uint32 buffer (:357)
        self._key_buffer = create_buffer(n_pixels, np.uint64)
        self._depth_buffer = DepthView(self._key_buffer, n_pixels)
        self._fill_mode = FillMode.WIREFRAME
        self._keys_armed = False
        self._gl_once = None    # draw_frame(): the (c_float * 48) the caller prepared, handed to the next draw as is
        self._owner = None      # set_scissor(): (c_int * 7) {x0, y0, x1, y1, stripe rows, mod, rem} handed to the native draws
        self._draws = None      # draws since the last clear(render_target): [(vertex buffer, its version, globals)], None = unknown

        assert len(vertex_shader.signature) == 2 and vertex_shader.return_annotation is not None, "Vertex shader signature incorrect. Must receive one argument with vertex type and another with globals type, and return another struct"
        assert len(fragment_shader.signature) == 2 and fragment_shader.return_annotation == float4, "Fragment shader signature incorrect. Must receive one argument with fragment type and another with globals type, and return a float4"
        self.vertex_input_type = vertex_shader.signature[0][1].annotation
        self.vertex_globals_type = vertex_shader.signature[1][1].annotation
        self.vertex_output_type = vertex_shader.return_annotation
        assert fragment_shader.signature[0][1].annotation == self.vertex_output_type, "Vertex shader output must be the same type than fragment shader input."
        self.fragment_globals_type = fragment_shader.signature[1][1].annotation
        self.vertex_shader = vertex_shader
        self.fragment_shader = fragment_shader
        self.vertex_shader_globals = vertex_shader_globals
        self.fragment_shader_globals = fragment_shader_globals

        self.shader_id = self._resolve_builtin()     # None -> user shaders: NVRTC-compiled generic pipeline
        self._generic = None if self.shader_id is not None else self._build_generic()
        self._scratch = None
        self._scratch_tris = -1
        # kept for API compatibility with code that reads them (:378-380); nothing is sized by them here
        self.fragments_capacity = 32 * render_target.width * render_target.height
        self.primitive_capacity = 200000

    # -- shader recognition -------------------------------------------------------------------------
    def _resolve_builtin(self):
        key = (shader_fingerprint(self.vertex_shader.source), shader_fingerprint(self.fragment_shader.source))
        sid = _BUILTIN_SHADERS.get(key)
        if sid is None or os.environ.get("RENDERTOY_B200_GENERIC_RASTER") == "1":
            return None
        f4, f3, f2, m4 = _core.float4, _core.float3, _core.float2, _core.float4x4
        try:
            _check_layout(self.vertex_input_type, [(f3, 0), (f3, 16), (f2, 32), (f3, 48), (f3, 64)], "vertex input")
            _check_layout(self.vertex_globals_type, [(m4, 0), (m4, 64), (m4, 128)], "vertex globals")
            if sid == _native.SHADER_LESSON08:
                _check_layout(self.vertex_output_type, [(f4, 0), (f3, 16)], "vertex output")
            else:
                _check_layout(self.vertex_output_type, [(f4, 0), (f3, 16), (f2, 32)], "vertex output")
                _check_layout(self.fragment_globals_type, [(_core.Texture2D, 0)], "fragment globals")
        except NotImplementedError:
            return None       # same shader text over different struct layouts: take the general path
        return sid

    def _build_generic(self):
        """User shaders: compile the general pipeline (rendering/raster_generic.cuh) around them with NVRTC."""
        from . import _dsl
        vout = np.dtype(self.vertex_output_type)
        first = vout.names[0]
        assert vout.fields[first][1] == 0 and vout.fields[first][0] == _core.float4, \
            "the first field of the vertex shader's output must be the float4 projected position"

        def all_float(dt):
            dt = np.dtype(dt)
            return all(all_float(dt.fields[n][0]) for n in dt.names) if dt.names else dt == np.float32
        assert all_float(vout), "every field of the vertex shader's output must be made of floats (they are interpolated)"
        src = _dsl.raster_program_source(self.vertex_shader, self.fragment_shader, self.vertex_input_type, vout,
                                         self.vertex_globals_type, self.fragment_globals_type)
        return {"module": _dsl.compile_program(src), "nf": vout.itemsize // 4, "dsl": _dsl, "rec": None}

    def _generic_draw(self, vertex_buffer, index_buffer, points):
        import ctypes
        g, rt = self._generic, self._render_target
        count = vertex_buffer.shape[0] if index_buffer is None else index_buffer.shape[0]
        n = count if points else count // 3
        assert vertex_buffer.dtype.itemsize == np.dtype(self.vertex_input_type).itemsize, "vertex buffer type differs from the vertex shader's input"
        rec_floats = n * g["nf"] * (1 if points else 6)
        if g["rec"] is None or g["rec"].size < rec_floats:
            g["rec"] = create_buffer(max(rec_floats, 4), np.float32)
        depth_bits = self._depth_buffer.take_pending()
        if not self._keys_armed:
            if depth_bits is None and int(self._key_buffer.version) == 0:
                depth_bits = 0
            self._keys_armed = True
        if depth_bits is not None:
            self._depth_buffer.fill(depth_bits)
        clear = rt.take_pending_clear()
        self._draws = None      # user vertex shader: no bound on where it puts the mesh
        clear_px = 0
        if clear is not None:
            v = np.clip(np.array([clear[2], clear[1], clear[0], clear[3]], np.float32) * np.float32(255.0), 0, 255)
            b = np.rint(v).astype(np.uint32)
            clear_px = int(b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24))
        P = ctypes.c_void_p
        ib = P(index_buffer.ptr) if index_buffer is not None else P(0)
        if index_buffer is not None:
            assert index_buffer.dtype == np.int32, "index buffer must be int32"
        W, H = np.int32(rt.width), np.int32(rt.height)
        kind = "points" if points else "triangles"
        g["dsl"].launch(g["module"], f"g_raster_{kind}", n,
                        [P(vertex_buffer.ptr), ib, self.vertex_shader_globals.get(), P(self._key_buffer.ptr), P(g["rec"].ptr), W, H])
        g["dsl"].launch(g["module"], f"g_resolve_{kind}", rt.width * rt.height,
                        [P(self._key_buffer.ptr), P(g["rec"].ptr), self.fragment_shader_globals.get(), P(rt.raw_ptr), W, H,
                         np.int32(0 if clear is None else 1), np.uint32(clear_px)])
        self._key_buffer.device_written()
        rt._buffer.device_written()

    # -- reference accessors (:385-397) ---------------------------------------------------------------
    def get_render_target(self):
        return self._render_target

    def get_depth_buffer(self):
        return self._depth_buffer

    @property
    def fill_mode(self) -> FillMode:
        return self._fill_mode

    @fill_mode.setter
    def fill_mode(self, value: FillMode):
        self._fill_mode = value

    # -- image-space partition (not in the reference, which is single-device: rendering/_core.py:10-11) ------------------
    def set_scissor(self, rect=None, stripes=None):
        """Restrict what the following draws (and the clears folded into them) may touch: the inclusive pixel rect
        (x0, y0, x1, y1) and / or the row stripes (rows, mod, rem) -- rows are grouped from y = 0 into stripes of `rows`
        rows and stripe s is owned iff s % mod == rem, which is how parallel.tile_rects deals one frame out to the ranks
        ("bin only into owned tiles").  Owned pixels (depth, winner, colour) come out exactly as without a
        scissor; everything else is left untouched.  set_scissor() with no arguments lifts it."""
        import ctypes
        if rect is None and stripes is None:
            self._owner = None
            return
        W, H = self._render_target.width, self._render_target.height
        x0, y0, x1, y1 = rect if rect is not None else (0, 0, W - 1, H - 1)
        rows, mod, rem = stripes if stripes is not None else (1, 1, 0)
        assert rows >= 1 and mod >= 1 and 0 <= rem < mod, "stripes = (rows per stripe, modulus, remainder)"
        assert self._generic is None, "a scissor / stripe partition needs the built-in (tutorial) shader pairs"
        self._owner = (ctypes.c_int * 7)(int(x0), int(y0), int(x1), int(y1), int(rows), int(mod), int(rem))

    @property
    def scissor(self):
        """None, or (x0, y0, x1, y1, stripe rows, stripe mod, stripe rem) as set by set_scissor()."""
        return None if self._owner is None else tuple(self._owner)

    # -- draws ------------------------------------------------------------------------------------------
    def draw_points(self, vertex_buffer, index_buffer=None):
        """Raster.draw_points (:399-414): one fragment per vertex (or per index), same depth / colour targets."""
        if self._generic is not None:
            return self._generic_draw(vertex_buffer, index_buffer, points=True)
        primitive_count = vertex_buffer.shape[0] if index_buffer is None else index_buffer.shape[0]
        pos4, nrm4 = _core.mesh_soa(vertex_buffer)
        idx_ptr = None
        if index_buffer is not None:
            assert index_buffer.dtype == np.int32, "index buffer must be int32 (_raster.py:142)"
            idx_ptr = index_buffer.ptr
        depth_bits = self._depth_buffer.take_pending()
        if not self._keys_armed:
            if depth_bits is None and int(self._key_buffer.version) == 0:
                depth_bits = 0
            self._keys_armed = True
        rt = self._render_target
        need = int(_native.lib().rt_raster_points_scratch_bytes(primitive_count))
        if self._scratch is None or self._scratch.nbytes < need:
            self._scratch = create_buffer(need, np.uint8)
            self._scratch_tris = -1        # forces draw_triangles to size its own layout next time
        gl, clear = self._vs_globals(), rt.take_pending_clear()
        self._note_draw(vertex_buffer, gl, clear is not None)
        _native.call("rt_raster_draw_points", pos4.data_ptr(), nrm4.data_ptr(), idx_ptr, primitive_count, self.shader_id,
                     gl, self._texture_handle(), rt.width, rt.height, self._key_buffer.ptr,
                     self._scratch.ptr, self._scratch.nbytes, rt.raw_ptr, clear,
                     0 if depth_bits is None else 1, depth_bits or 0, self._owner, stream_ptr())
        self._key_buffer.device_written()
        rt._buffer.device_written()

    def _note_draw(self, vertex_buffer, globals48, new_frame):
        """Book-keeping for content_rect (no device work, no sync): which mesh went in under which transforms since the
        render target was last cleared."""
        if new_frame:
            self._draws = []
        if self._draws is not None:
            if len(self._draws) >= 64:
                self._draws = None
            else:
                self._draws.append((vertex_buffer, vertex_buffer.version, globals48))

    @property
    def content_rect(self):
        """Inclusive pixel rect (x0, y0, x1, y1) outside of which the render target holds the colour of the last
        clear(render_target): the union, over the draws since that clear, of the screen rectangles of the bounding boxes
        of 64 chunks of each mesh under the draw's World/View/Proj (rt_raster_screen_bounds_n).  The whole frame when that is not known: user
        vertex shaders, a mesh reaching the near plane, a vertex buffer modified since it was drawn, no clear yet.
        (x1 < x0: nothing was drawn on screen.)  Not part of the reference API; it is what lets a frame be read back or
        gathered sparsely (parallel.SparseFrameCopier).  The first query for a mesh version reads its bounds back (a sync)."""
        import ctypes
        W, H = self._render_target.width, self._render_target.height
        full = (0, 0, W - 1, H - 1) if self._owner is None else \
            (max(0, self._owner[0]), max(0, self._owner[1]), min(W - 1, self._owner[2]), min(H - 1, self._owner[3]))
        if self._draws is None:
            return full
        x0, y0, x1, y1 = W, H, -1, -1
        r = (ctypes.c_int * 4)()
        for vb, version, gl in self._draws:
            if vb.version != version:
                return full
            lo, hi, k = _core.mesh_chunk_bounds(vb)
            if not _native.lib().rt_raster_screen_bounds_n(gl, lo, hi, k, W, H, r):
                return full
            if r[2] >= r[0] and r[3] >= r[1]:
                x0, y0, x1, y1 = min(x0, r[0]), min(y0, r[1]), max(x1, r[2]), max(y1, r[3])
        if self._owner is not None:     # nothing outside the scissor rect was written
            x0, y0, x1, y1 = max(x0, self._owner[0]), max(y0, self._owner[1]), min(x1, self._owner[2]), min(y1, self._owner[3])
        return (x0, y0, x1, y1) if (x1 >= x0 and y1 >= y0) else (0, 0, -1, -1)

    def _vs_globals(self):
        """48 floats (World, View, Proj) as a ctypes array: the Transforms struct is three contiguous float4x4 (layout
        checked in _resolve_builtin), read straight from the struct's host shadow."""
        return _native.float_array_from_bytes(self.vertex_shader_globals.host_bytes(), 48)

    def _texture_handle(self):
        if self.shader_id != _native.SHADER_LESSON09:
            return 0
        g = self.fragment_shader_globals.get()
        desc = g[g.dtype.names[0]]
        return _core.__MEMORY_POOL__.texture_handle(int(desc["offset"]))

    def draw_frame(self, vertex_buffer, index_buffer=None, transforms=None, clear_color=0.0, clear_depth=1.0):
        """One tutorial frame in ONE call (not in the reference API): what lesson08:90-105 spells as

            with mapped(shader_globals) as map: map["World"], map["View"], map["Proj"] = ...
            clear(raster.get_render_target()); clear(raster.get_depth_buffer(), 1.0); raster.draw_triangles(vb, ib)

        transforms: None (keep what the globals buffer holds), a (World, View, Proj) tuple of float4x4 values, or 48 floats /
        a (c_float * 48) prepared once per camera.  The globals buffer is updated, so the four calls above and this one are
        interchangeable frame by frame; the result is bit-identical.  It exists because the four calls cost 35-40 us of
        Python per frame while the B200 needs ~50 us for the frame itself.  clear_color: a float or
        None (no clear); clear_depth: a float or None."""
        import ctypes
        assert self._generic is None, "draw_frame needs the built-in (tutorial) shader pairs"
        if transforms is not None:
            if isinstance(transforms, tuple):
                transforms = transforms48(*transforms)
            g = self.vertex_shader_globals
            st = g._st
            if st.host is not None:      # straight into the struct's host shadow (what mapped() would do), no device traffic
                if isinstance(transforms, ctypes.Array):
                    ctypes.memmove(st.host.ctypes.data + g.offset, transforms, 192)
                    self._gl_once = transforms
                else:
                    st.host[g.offset:g.offset + 192] = np.ascontiguousarray(transforms, np.float32).view(np.uint8)
                st.host_valid, st.dev_valid = True, False
                st.version += 1
            else:
                g.set(np.ascontiguousarray(transforms, np.float32).view(g.dtype).reshape(()))
        if clear_color is not None:
            self._render_target._pending_clear = _native.float4_const(float(clear_color))
        if clear_depth is not None:
            self._depth_buffer._pending = int(np.float32(clear_depth).view(np.uint32))
        self.draw_triangles(vertex_buffer, index_buffer)

    def draw_triangles(self, vertex_buffer, index_buffer):
        """Raster.draw_triangles (:416-437).  index_buffer None -> triangle soup; else int32 indices.
        Accumulates into the persistent depth / colour targets exactly like consecutive reference draws."""
        if self._generic is not None:
            return self._generic_draw(vertex_buffer, index_buffer, points=False)
        primitive_count = (vertex_buffer.shape[0] if index_buffer is None else index_buffer.shape[0]) // 3
        pos4, nrm4 = _core.mesh_soa(vertex_buffer)
        idx_ptr = None
        if index_buffer is not None:
            assert index_buffer.dtype == np.int32, "index buffer must be int32 (_raster.py:154)"
            idx_ptr = index_buffer.ptr
        depth_bits = self._depth_buffer.take_pending()     # a deferred clear(depth_buffer, v) rides along with this draw
        if not self._keys_armed:
            # first draw on a never-cleared target: give the zero-filled key buffer its NO_PRIMITIVE low words
            if depth_bits is None and int(self._key_buffer.version) == 0:
                depth_bits = 0
            self._keys_armed = True
        rt = self._render_target
        if primitive_count > self._scratch_tris:
            need = _native.lib().rt_raster_scratch_bytes(self.shader_id, primitive_count, rt.width, rt.height)
            self._scratch = create_buffer(int(need), np.uint8)   # zero-filled: the control block starts armed
            self._scratch_tris = primitive_count
        gl, clear = self._gl_once or self._vs_globals(), rt.take_pending_clear()
        self._gl_once = None
        self._note_draw(vertex_buffer, gl, clear is not None)
        _native.call("rt_raster_draw_triangles", pos4.data_ptr(), nrm4.data_ptr(), idx_ptr, primitive_count, self.shader_id,
                     gl, self._texture_handle(), rt.width, rt.height, self._key_buffer.ptr,
                     self._scratch.ptr, self._scratch.nbytes, rt.raw_ptr, clear,
                     0 if depth_bits is None else 1, depth_bits or 0, self._owner, stream_ptr())
        self._key_buffer.device_written()
        rt._buffer.device_written()
