"""Host-side logic of the `rendering` mirror (no GPU): dtype layouts, matrices, buffers, loaders, shader
recognition, and the C ABI's symbol table."""
import os
import re
import subprocess

import numpy as np
import pytest

from rendertoy_b200 import _native, lessons, scenes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_public_names_match_reference_init(ren):
    names = """kernel_main kernel_struct create_buffer_from create_buffer create_struct mapped kernel_function create_struct_from
        create_image2d Image Buffer r_image1d_t w_image1d_t r_image2d_t w_image2d_t r_image3d_t w_image3d_t make_float2 make_float3
        make_float4 make_float4x4 translate identity scale rotate matmul to_array clear perspective look_at normalize dot float2
        float3 float4 int2 int3 int4 uint2 uint3 uint4 float4x4 Texture2D create_texture2D Mesh WeldMode SubdivisionMode MeshVertex
        manifold load_obj create_presenter Presenter Event Raster""".split()
    for n in names:                                   # rendering/__init__.py:1-15
        assert hasattr(ren, n), n
    import rendering._raster, rendering._core, rendering._raycaster   # import-by-path used by reference callers
    assert rendering._raster.FillMode.SOLID == 3 and hasattr(rendering._core, "make_int2") and hasattr(rendering._core, "cross")


def test_struct_layouts(ren):
    """SURVEY.md appendix B (OpenCL layout rules)."""
    mv = ren.MeshVertex
    assert mv.itemsize == 80 and [mv.fields[n][1] for n in "PNCTB"] == [0, 16, 32, 48, 64]
    assert ren.float3.itemsize == 16 and ren.float4x4.itemsize == 64 and ren.float2.itemsize == 8
    assert ren.Texture2D.itemsize == 12

    @ren.kernel_struct
    class Transforms:
        World: ren.float4x4
        View: ren.float4x4
        Proj: ren.float4x4

    @ren.kernel_struct
    class VO8:
        proj: ren.float4
        C: ren.float3

    @ren.kernel_struct
    class VO9:
        proj: ren.float4
        L: ren.float3
        C: ren.float2

    assert Transforms.itemsize == 192 and VO8.itemsize == 32 and VO9.itemsize == 48
    assert VO9.fields["C"][1] == 32
    from rendering._raycaster import BVH_AABB, BVH_Triangle
    assert BVH_AABB.itemsize == 48 and BVH_Triangle.itemsize == 64      # `int` -> int64 on Linux


def test_matrices_bitwise_equal_oracle_restatement(ren, oracle):
    hm = oracle.host_math
    f3 = ren.make_float3
    for eye in ((0, 0.3, 1.0), (0, 0.3, 2), (1.5, -0.7, 0.4)):
        a = ren.to_array(ren.look_at(f3(*eye), f3(0, 0, 0), f3(0, 1, 0)))
        assert np.array_equal(a.view(np.uint32), hm.look_at(eye, (0, 0, 0), (0, 1, 0)).view(np.uint32))
    for asp in (1.0, 640 / 480, 1920 / 1080, 3840 / 2160):
        assert np.array_equal(ren.to_array(ren.perspective(aspect_ratio=asp)).view(np.uint32), hm.perspective(aspect_ratio=asp).view(np.uint32))
    for t in (0.0, 0.5, 2.1, 6.0):
        w = np.array(ren.matmul(ren.scale(1.0), ren.rotate(t, f3(0, 1, 0))), dtype=ren.float4x4)
        assert np.array_equal(ren.to_array(w).view(np.uint32), hm.matmul(hm.scale(1.0), hm.rotate(t, (0, 1, 0))).view(np.uint32))
    ax = ren.normalize(f3(1, 2, 3))
    assert np.array_equal(ren.to_array(ren.rotate(0.7, ax)).view(np.uint32), hm.rotate(0.7, ren.to_array(ax)).view(np.uint32))
    assert isinstance(ren.matmul(ren.identity(), ren.identity()), tuple)     # reference returns .item() tuples here
    assert ren.dot(f3(1, 2, 3), f3(4, 5, 6)) == 32.0
    assert np.allclose(ren.to_array(ren.translate(1, 2, 3))[3, :3], (1, 2, 3))


def test_buffers_structs_clear_and_views(ren):
    b = ren.create_buffer(10, np.float32)
    assert b.shape == (10,) and (b.get() == 0).all()
    with ren.mapped(b) as m:
        m[:] = np.arange(10)
    assert (b.get() == np.arange(10)).all() and (b[2:5].get() == [2, 3, 4]).all() and len(b) == 10
    ren.clear(b, 1.0)
    assert (b.get().view(np.uint32) == 0x3F800000).all()
    s = ren.create_struct(ren.Texture2D)
    with ren.mapped(s) as m:
        m["width"] = 7
    assert s.shape == () and int(s.get()["width"]) == 7
    big = ren.create_buffer(100_000, ren.float4)            # no host shadow at this size
    with ren.mapped(big) as m:
        m["x"] = 3.0
    assert (big.get()["x"] == 3.0).all() and big.view(np.float32).shape == (400_000,)
    c = ren.create_buffer_from(np.arange(6, dtype=np.int32))
    assert int(c[5:6].map_to_host()[0]) == 5
    mem, desc = ren.create_texture2D(5, 3)
    assert mem.shape == (3, 5) and mem.dtype == ren.float4 and int(desc.get()["width"]) == 5 and int(desc.get()["height"]) == 3
    _, desc2 = ren.create_texture2D(2, 2)
    assert int(desc2.get()["offset"]) % 512 == 0 and int(desc2.get()["offset"]) >= 5 * 3 * 16


def test_manifold(ren):
    m = ren.manifold(4, 3)
    v = m.vertices.get()
    assert m.vertices.shape == (20,) and m.indices.shape == (4 * 3 * 6,) and m.indices.dtype == np.int32
    rows = v.view(np.float32).reshape(-1, 20)
    assert np.allclose(rows[6, 0:3], (0.25, 1 / 3, 0.0)) and np.allclose(rows[:, 8:10], rows[:, 0:2])
    idx = m.indices.get()
    assert list(idx[:6]) == [0, 1, 5, 1, 2, 6] and list(idx[12:18]) == [0, 5, 4, 1, 6, 5]      # row stride `slices` quirk kept
    with pytest.raises(Exception, match="Not implemented yet"):
        m.clone()


def test_load_obj_roundtrip(ren, tmp_path):
    rows = scenes.dragon(300, normalise=False)
    path = tmp_path / "mini.obj"
    scenes.write_obj(str(path), rows, with_uv=True)
    (mesh, material), = ren.load_obj(str(path))
    assert material is None and mesh.vertices.shape == (rows.shape[0],)
    got = mesh.vertices.get().view(np.float32).reshape(-1, 20)
    want = scenes.normalise_like_load_obj(np.loadtxt([f"{r[0]:.9g} {r[1]:.9g} {r[2]:.9g}" for r in rows], dtype=np.float32).reshape(-1, 3).copy())
    assert np.array_equal(got[:, 0:3], want[:, 0:3]) if want.shape[1] >= 3 else True
    assert np.allclose(got[:, 4:7], rows[:, 4:7], atol=1e-6) and np.allclose(got[:, 8:10], rows[:, 8:10], atol=1e-6)
    assert got[:, 0:3].min() == -0.5 and mesh.indices.shape == (rows.shape[0],) and (mesh.indices.get() == 0).all()
    # quads fan-triangulate, negative indices resolve
    q = tmp_path / "quad.obj"
    q.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvn 0 0 1\nf -4//1 -3//1 -2//1 -1//1\n")
    (mesh2, _), = ren.load_obj(str(q))
    p = mesh2.vertices.get().view(np.float32).reshape(-1, 20)[:, 0:3] + 0.5
    assert p.shape[0] == 6 and np.allclose(p[[0, 1, 2, 3, 4, 5]], [(0, 0, 0), (1, 0, 0), (1, 1, 0), (0, 0, 0), (1, 1, 0), (0, 1, 0)])


def test_raster_recognises_tutorial_shaders_and_rejects_others(ren):
    pres = ren.create_presenter(32, 16)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    assert raster.shader_id == _native.SHADER_LESSON08 and raster.get_depth_buffer().shape == (512,)
    tex = np.zeros((4, 4, 3), np.uint8)
    raster9, *_ = lessons.build_lesson09(ren, pres.get_render_target(), tex)
    assert raster9.shader_id == _native.SHADER_LESSON09

    @ren.kernel_struct
    class VOut:
        proj: ren.float4
        C: ren.float3

    @ren.kernel_function
    def vs(vertex: ren.MeshVertex, info: ren.float4x4) -> VOut:
        """
        VOut o; o.proj = (float4)(vertex.P, 1); o.C = vertex.N; return o;
        """

    @ren.kernel_function
    def fs(fragment: VOut, info: ren.float4x4) -> ren.float4:
        """
        return (float4)(fragment.C, 1);
        """

    # a user-written pair is compiled with NVRTC around the general pipeline (works without a GPU; drawing needs one)
    custom = ren.Raster(pres.get_render_target(), vs, ren.create_struct(ren.float4x4), fs, ren.create_struct(ren.float4x4))
    assert custom.shader_id is None and custom._generic["nf"] == 8

    @ren.kernel_function
    def broken_fs(fragment: VOut, info: ren.float4x4) -> ren.float4:
        """
        return not_a_function(fragment.C);
        """

    with pytest.raises(RuntimeError, match="failed to build"):
        ren.Raster(pres.get_render_target(), vs, ren.create_struct(ren.float4x4), broken_fs, ren.create_struct(ren.float4x4))
    with pytest.raises(AssertionError, match="Fragment shader signature incorrect"):
        ren.Raster(pres.get_render_target(), vs, None, vs, None)
    with pytest.raises(Exception, match="Can not call to this function from host"):
        vs()


def test_presenter_is_headless_and_bounded(ren, monkeypatch):
    monkeypatch.setenv("RENDERTOY_B200_FRAMES", "2")
    p = ren.create_presenter(8, 4)
    n = 0
    while True:
        ev, _ = p.poll_events()
        if ev == ren.Event.CLOSED:
            break
        p.present()
        n += 1
    assert n == 2 and p.get_render_target().shape == (8, 4)


def test_c_abi_exports_every_declared_symbol():
    """include/rendertoy_b200.h <-> librendertoy_b200.so <-> ctypes table."""
    header = open(os.path.join(ROOT, "include", "rendertoy_b200.h")).read()
    declared = set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", header))
    declared -= {"rt_status"}
    lib = _native.lib()                      # loads without a GPU
    exported = set(re.findall(r" T (rt_\w+)", subprocess.run(["nm", "-D", "--defined-only", _native.LIB_PATH], capture_output=True, text=True).stdout))
    assert declared, "no declarations parsed"
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert declared == set(_native.SIGNATURES), f"ctypes table out of sync: {sorted(declared ^ set(_native.SIGNATURES))}"
    assert lib.rt_abi_version() == 1
    assert lib.rt_raster_scratch_bytes(8, 1000, 64, 64) > 2 * 1000 * 64 and lib.rt_bvh_node_bytes(1000) == 999 * 64


def test_kernels_fail_loudly_without_cuda(ren):
    import torch
    if torch.cuda.is_available():
        pytest.skip("has a GPU")
    pres = ren.create_presenter(32, 16)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    vb = ren.create_buffer(3, ren.MeshVertex)
    with pytest.raises(Exception):
        raster.draw_triangles(vb, None)       # no CPU fallback: the native call refuses


def test_scalar_rotate_and_look_at_equal_the_numpy_formulations(ren):
    """rotate()/look_at() run on Python scalars with explicit float32 roundings (a frame calls them once each and the
    array versions cost ~90 us together); they must reproduce the NumPy formulations bit for bit, degenerate input included."""
    import warnings
    from rendering import _core
    rng = np.random.default_rng(11)

    def f3(scale=1.0):
        return ren.make_float3(*(rng.standard_normal(3) * scale).astype(np.float32).tolist())

    for k in range(1500):
        axis = f3(10.0 ** rng.integers(-6, 6))
        angle = float(rng.uniform(-10, 10)) if k % 3 else np.float32(rng.uniform(-10, 10))
        assert np.asarray(_core.rotate(angle, axis)).tobytes() == np.asarray(_core._rotate_numpy(angle, axis)).tobytes()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for k in range(1500):
            s = 10.0 ** rng.integers(-3, 4)
            cam, tgt, up = f3(s), f3(s), f3()
            if k % 5 == 0:
                up = ren.make_float3(0, 1, 0)
            if k % 11 == 0:
                tgt = ren.make_float3(0, 0, 0)
            if k % 300 == 0:
                tgt = cam                                   # zero view direction: NumPy's nan pattern must come through
            if k % 301 == 0:
                up = ren.make_float3(*[float(x) for x in (ren.to_array(tgt) - ren.to_array(cam))])   # up parallel to the view direction
            a, b = _core.look_at(cam, tgt, up), _core._look_at_numpy(cam, tgt, up)
            assert np.asarray(a).tobytes() == np.asarray(b).tobytes()


def test_camera_frame_native_matches_float64_numpy(ren):
    """rt_camera_frame (host arithmetic in the native library) against the float64 NumPy formulation: same frame up to the
    last float32 bit of components that are not cancellation noise; singular input raises."""
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes

    def numpy_frame(view, proj, world):
        m = lambda x: ren.to_array(np.asarray(x, dtype=ren.float4x4)).astype(np.float64).reshape(4, 4)
        view, proj = m(view), m(proj)
        r = view[0:3, 0:3]
        eye = -(view[3, 0:3] @ np.linalg.inv(r))
        u, v, w = r[:, 0] / proj[0, 0], r[:, 1] / proj[1, 1], r[:, 2]
        if world is not None:
            winv = np.linalg.inv(m(world))
            eye = (np.append(eye, 1.0) @ winv)[:3]
            u, v, w = ((np.append(x, 0.0) @ winv)[:3] for x in (u, v, w))
        return np.concatenate([eye, u, v, w])

    for k in range(200):
        world, view, proj = scenes.lesson_camera(ren, 6 if k % 2 else 8, 0.0731 * k, 1920, 1080)
        if k % 4 == 0:      # scaled, translated world: the general inverse
            world = ren.matmul(ren.scale(1.0 + 0.01 * k, 0.5, 2.0), np.array(ren.translate(0.1, 0.02 * k, -0.3), dtype=ren.float4x4))
        world = None if k % 7 == 0 else np.array(world, dtype=ren.float4x4)
        got = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), world)
        want = numpy_frame(view, proj, world)
        assert got.dtype == np.float32 and got.shape == (12,)
        assert np.allclose(got, want, rtol=3e-7, atol=1e-7 * np.abs(want).max())
    with pytest.raises(np.linalg.LinAlgError):
        camera_frame(np.zeros((4, 4), np.float32), np.eye(4, dtype=np.float32))


_TRICKY_OBJ = """# comment line
mtllib nothing.mtl
v 0 0 0
v 1 0 0 0.5 0.5 0.5
v 1 1 0
v 0 1 0
v 0.5 0.5 1e0
v -.25 +2.5e-1 1.
v 12345678901234567890.5 1e-30 0.1234567890123456789
vn 0 0 1
vn 0 1 0
vn 1 0 0
vt 0 0
vt 1 0 0
vt 1 1
vt 0.25
f 1 2 3
usemtl first
o quad
f 1/1/1 2/2/1 3/3/2 4/4/3
f -4//-1 -3//-2 -2//-3
g ignored group
s off
o second
f 5/1/1 6/2/2 7/3/3 1/4/1 2/1/2
usemtl other
f 1 2 3
o third
usemtl first
f 2/2/2 3/3/3 4/4/1
o empty_one
\tv 9 9 9
   # indented comment
f 8/1/1 1/1/1 2/2/2\r
"""


def _compare_loaders(ren, path):
    from rendering import _loaders
    a, b = ren.load_obj(str(path)), _loaders._load_obj_python(str(path))
    assert len(a) == len(b)
    for (ma, _), (mb, _) in zip(a, b):
        ra, rb = ma.vertices.get().view(np.float32).reshape(-1, 20), mb.vertices.get().view(np.float32).reshape(-1, 20)
        assert ra.shape == rb.shape
        assert np.array_equal(ra.view(np.uint32), rb.view(np.uint32)), "native OBJ parser differs from the Python formulation"
        assert ma.indices.shape == mb.indices.shape and ma.indices.dtype == mb.indices.dtype
    return a


def test_native_obj_parser_matches_python_formulation(ren, tmp_path):
    p = tmp_path / "tricky.obj"
    p.write_bytes(_TRICKY_OBJ.encode())
    objs = _compare_loaders(ren, p)
    # the face before any `o` makes an anonymous mesh on material default0; the four named meshes all list material `first`
    # first, whose corner list is shared between them (quad + triangle + pentagon fan + triangle + triangle = 24 corners)
    assert [(m.vertices.shape[0], m.indices.shape[0]) for m, _ in objs] == [(3, 3), (24, 9), (24, 12), (24, 3), (24, 3)]
    # the dragon stand-in through both parsers, with and without UVs, and with CRLF line ends
    rows = scenes.dragon(700)
    for k, with_uv in enumerate((False, True)):
        q = tmp_path / f"dragon{k}.obj"
        scenes.write_obj(str(q), rows, with_uv=with_uv)
        _compare_loaders(ren, q)
    crlf = tmp_path / "crlf.obj"
    crlf.write_bytes((tmp_path / "dragon1.obj").read_bytes().replace(b"\n", b"\r\n"))
    _compare_loaders(ren, crlf)
    # malformed input raises, with the line number in the message
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 zero\nf 1 2 3\n")
    with pytest.raises(Exception, match="line 3"):
        ren.load_obj(str(bad))
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 9\n")
    with pytest.raises(Exception, match="out of range"):
        ren.load_obj(str(bad))
    with pytest.raises(Exception, match="cannot open"):
        ren.load_obj(str(tmp_path / "missing.obj"))
    empty = tmp_path / "empty.obj"
    empty.write_text("")
    assert ren.load_obj(str(empty)) == []


def test_raster_screen_bounds_native(ren):
    """rt_raster_screen_bounds (host-only C): rectangle of the projected mesh box against a float64 NumPy projection of the
    same corners; no bound when the box reaches the near plane."""
    import ctypes
    L = _native.lib()
    lo, hi = np.array([-0.5, -0.31, -0.22]), np.array([0.5, 0.27, 0.24])
    clo, chi = (ctypes.c_double * 3)(*lo), (ctypes.c_double * 3)(*hi)
    corners = np.array([[x, y, z, 1.0] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
    for k in range(48):
        W, H = [(1920, 1080), (333, 211)][k % 2]
        mats = [ren.to_array(np.array(m, dtype=ren.float4x4)).astype(np.float32) for m in scenes.lesson_camera(ren, 8 if k % 3 else 6, 0.41 * k, W, H)]
        g = np.concatenate([m.ravel() for m in mats]).astype(np.float32)
        rect = (ctypes.c_int * 4)()
        ok = L.rt_raster_screen_bounds(g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), clo, chi, W, H, rect)
        h4 = corners @ mats[0].astype(np.float64) @ mats[1].astype(np.float64) @ mats[2].astype(np.float64)
        assert ok == 1 and (h4[:, 2] > 0).all()
        px = (h4[:, 0] / h4[:, 3] + 1.0) * W / 2
        py = (1.0 - h4[:, 1] / h4[:, 3]) * H / 2
        want = (max(0, int(np.floor(px.min())) - 2), max(0, int(np.floor(py.min())) - 2),
                min(W - 1, int(np.floor(px.max())) + 3), min(H - 1, int(np.floor(py.max())) + 3))
        assert tuple(rect) == want
    from oracle import host_math as hm
    g = np.concatenate([m.ravel() for m in (hm.rotate(4.5, (0, 1, 0)), hm.look_at((0.12, 0.32, 0.3), (0, 0, 0), (0, 1, 0)),
                                            hm.perspective(aspect_ratio=16 / 9))]).astype(np.float32)
    assert L.rt_raster_screen_bounds(g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), clo, chi, 640, 360, (ctypes.c_int * 4)()) == 0
    # n boxes (Raster.content_rect: 64 chunks of the mesh): the union of the single rectangles; no bound if one box has none
    rng = np.random.default_rng(8)
    centres = rng.uniform(-0.4, 0.4, (9, 3))
    blo, bhi = centres - rng.uniform(0.01, 0.08, (9, 3)), centres + rng.uniform(0.01, 0.08, (9, 3))
    nlo, nhi = (ctypes.c_double * 27)(*blo.ravel()), (ctypes.c_double * 27)(*bhi.ravel())
    for k in range(12):
        W, H = 1920, 1080
        mats = [ren.to_array(np.array(m, dtype=ren.float4x4)).astype(np.float32) for m in scenes.lesson_camera(ren, 8 if k % 2 else 6, 0.5 * k, W, H)]
        gp = np.concatenate([m.ravel() for m in mats]).astype(np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
        singles = []
        for i in range(9):
            r1 = (ctypes.c_int * 4)()
            assert L.rt_raster_screen_bounds(gp, (ctypes.c_double * 3)(*blo[i]), (ctypes.c_double * 3)(*bhi[i]), W, H, r1) == 1
            if r1[2] >= r1[0] and r1[3] >= r1[1]:
                singles.append(tuple(r1))
        rn = (ctypes.c_int * 4)()
        assert L.rt_raster_screen_bounds_n(gp, nlo, nhi, 9, W, H, rn) == 1
        assert tuple(rn) == (min(r_[0] for r_ in singles), min(r_[1] for r_ in singles), max(r_[2] for r_ in singles), max(r_[3] for r_ in singles))
    assert L.rt_raster_screen_bounds_n(g.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), nlo, nhi, 9, 640, 360, (ctypes.c_int * 4)()) == 0


def _random_obj(rng):
    """A random but valid OBJ text touching every rule of the loader: all four corner formats (fixed per material by its
    first corner, later faces may differ), negative indices, polygons, several `o` / `usemtl`, ignored statements, odd
    number spellings, tabs, trailing blanks, CRLF."""
    def num():
        x = float(rng.normal()) * 10.0 ** int(rng.integers(-3, 4))
        style = int(rng.integers(0, 6))
        return [f"{x:.9g}", f"{x:.3f}", f"{x:e}", f"{x:+.5g}", repr(x), f"{int(x)}"][style]
    lines, nv, nt, nn = [], 0, 0, 0
    eol = "\r\n" if rng.random() < 0.3 else "\n"
    sep = lambda: " " * int(rng.integers(1, 3)) if rng.random() < 0.8 else "\t"
    def emit(*toks):
        lines.append(sep().join(toks) + (" " if rng.random() < 0.1 else ""))
    for _ in range(int(rng.integers(3, 9))):
        emit("v", num(), num(), num()); nv += 1
    emit("vt", num(), num()); nt += 1
    emit("vn", num(), num(), num()); nn += 1
    for _ in range(int(rng.integers(4, 40))):
        kind = rng.random()
        if kind < 0.25:
            emit("v", num(), num(), num(), *([num()] if rng.random() < 0.2 else [])); nv += 1
        elif kind < 0.35:
            emit("vt", num(), *([num()] if rng.random() < 0.8 else [])); nt += 1
        elif kind < 0.45:
            emit("vn", num(), num(), num()); nn += 1
        elif kind < 0.52:
            emit("o", f"mesh{int(rng.integers(0, 5))}")
        elif kind < 0.60:
            emit("usemtl", *([f"mat{int(rng.integers(0, 4))}"] if rng.random() < 0.9 else []))
        elif kind < 0.68:
            emit(str(rng.choice(["g grp", "s 1", "mtllib x.mtl", "# note", "l 1 2", ""])))
        else:
            fmt = int(rng.integers(0, 4))
            corners = []
            for _ in range(int(rng.integers(3, 7))):
                def idx(count):
                    return str(int(rng.integers(1, count + 1))) if rng.random() < 0.7 else str(-int(rng.integers(1, count + 1)))
                v, t, n = idx(nv), idx(nt), idx(nn)
                corners.append([v, f"{v}/{t}", f"{v}//{n}", f"{v}/{t}/{n}"][fmt])
            emit("f", *corners)
    return eol.join(lines) + (eol if rng.random() < 0.8 else "")


def test_native_obj_parser_fuzz(ren, tmp_path):
    rng = np.random.default_rng(2024)
    p = tmp_path / "fuzz.obj"
    meshes = 0
    for _ in range(150):
        p.write_bytes(_random_obj(rng).encode())
        meshes += len(_compare_loaders(ren, p))
    assert meshes > 100


def test_host_math_equals_the_reference_run(ren):
    """tests/golden/host_math_reference.npz holds outputs of the reference's OWN scale / translate / identity / rotate /
    perspective / matmul / dot (tests/golden/make_host_math_golden.py runs them with a stand-in for pyopencl); the package and
    the oracle's host_math must reproduce them bit for bit.  (look_at / normalize raise in the reference under NumPy 2 and are
    pinned by SURVEY appendix D's known answers only.)"""
    from oracle import host_math as hm
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_math_reference.npz"))

    def flat(m):
        return np.asarray(m if not isinstance(m, tuple) else np.array(m, dtype=ren.float4x4)).reshape(-1).view(np.float32)[:16].copy()

    def same(got, want):
        return np.array_equal(np.asarray(got, np.float32).view(np.uint32), np.asarray(want, np.float32).view(np.uint32))

    f3 = lambda v: ren.make_float3(*[float(x) for x in v])
    assert same(np.stack([flat(ren.rotate(float(a), f3(ax))) for a, ax in zip(z["rotate_angle"], z["rotate_axis"])]), z["rotate"])
    assert same(np.stack([flat(ren.rotate(np.float32(a), f3(ax))) for a, ax in zip(z["rotate_angle"], z["rotate_axis"])]), z["rotate_f32_angle"])
    assert same(np.stack([flat(ren._core._rotate_numpy(np.float32(a), f3(ax))) for a, ax in zip(z["rotate_angle"], z["rotate_axis"])]), z["rotate_f32_angle"])
    assert same(np.stack([hm.rotate(float(a), tuple(float(x) for x in ax)).ravel() for a, ax in zip(z["rotate_angle"], z["rotate_axis"])]), z["rotate"])
    assert same(np.stack([flat(ren.scale(*[float(x) for x in s])) for s in z["scale_in"]]), z["scale"])
    assert same(np.stack([flat(ren.scale(float(s[0]))) for s in z["scale_in"]]), z["scale_uniform"])
    assert same(np.stack([flat(ren.translate(*[float(x) for x in t])) for t in z["translate_in"]]), z["translate"])
    assert same(np.stack([flat(ren.translate(f3(t))) for t in z["translate_in"]]), z["translate_vec"])
    assert same(flat(ren.identity()), z["identity"])
    assert same(np.stack([flat(ren.perspective(f, a, n, fa)) for f, a, n, fa in z["perspective_in"]]), z["perspective"])
    assert same(np.stack([hm.perspective(f, a, n, fa).ravel() for f, a, n, fa in z["perspective_in"]]), z["perspective"])
    assert same(np.stack([flat(ren.perspective(aspect_ratio=w / h)) for w, h in ((512, 512), (1920, 1080), (3840, 2160), (333, 211))]),
                z["perspective_default_aspect"])
    m44 = lambda v: ren.make_float4x4(*[float(x) for x in v])
    assert same(np.stack([flat(ren.matmul(m44(a), m44(b))) for a, b in zip(z["matmul_a"], z["matmul_b"])]), z["matmul"])
    vec = lambda v, b: np.asarray(np.array(ren.matmul(ren.make_float4(*[float(x) for x in v]), m44(b)), dtype=ren.float4)).reshape(-1).view(np.float32)[:4]
    assert same(np.stack([vec(v, b) for v, b in zip(z["matmul_vec_a"], z["matmul_b"])]), z["matmul_vec"])
    assert same(np.array([ren.dot(f3(a), f3(b)) for a, b in zip(z["dot_a"], z["dot_b"])]), z["dot"])
    assert same(np.stack([flat(ren.matmul(ren.scale(1.0), ren.rotate(float(t), ren.make_float3(0, 1, 0)))) for t in z["world_t"]]), z["world"])
    assert same(np.stack([hm.matmul(hm.scale(1.0), hm.rotate(float(t), (0, 1, 0))).ravel() for t in z["world_t"]]), z["world"])


def test_every_image_format_can_be_created_with_either_dtype_spelling(ren):
    """create_image2d with each key of get_valid_image_formats(), and with the np.float32 CLASS the reference's table is
    keyed by (rendering/_core.py:343-349)."""
    from rendering import _core
    for fmt in _core.get_valid_image_formats():
        im = ren.create_image2d(6, 4, fmt)
        assert im.shape == (6, 4)
        with ren.mapped(im) as m:
            assert m.shape[:2] == (4, 6)
    im = ren.create_image2d(6, 4, np.float32)
    with ren.mapped(im) as m:
        assert m.shape == (4, 6) and m.dtype == np.float32


def test_image_on_adopted_memory_reads_that_memory(ren):
    """Image(..., memory=tensor) (a FrameStore slot): reads see what the memory holds, also after somebody else wrote it."""
    import torch
    from rendering import _core
    mem = torch.full((8 * 4 * 4,), 9, dtype=torch.uint8, device=_core.device())
    im = ren.Image(8, 4, _core.RGBA, memory=mem)
    assert int(im.get()[0, 0, 0]) == 9
    mem.fill_(5)        # an external writer (rt_copy_rect, a peer kernel)
    assert int(im.get()[3, 7, 3]) == 5
    with ren.mapped(im) as m:
        m[0, 0, 0] = 1
    assert int(mem[0]) == 1 and int(mem[1]) == 5


def test_raster_refuses_a_render_target_it_cannot_write(ren):
    from rendertoy_b200 import lessons
    with pytest.raises(AssertionError):
        lessons.build_lesson08(ren, ren.create_image2d(8, 8, ren.float4))


def test_bench_reference_arm_prints_the_contract_line():
    """bench.py --impl reference: one JSON line on stdout with the contract's keys, the SAME `config` dict our arm prints, all
    host cores used whatever OMP_NUM_THREADS says (torchrun exports 1), and nothing under rendertoy_b200/ imports the oracle."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--path", "raster", "--steps", "2", "--warmup", "0"],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    sys.path.insert(0, root)
    import bench
    assert d["impl"] == "reference" and d["config"] == bench.CONFIG_RAS and d["metric"] == bench.METRIC_RAS and d["unit"] == "Mtris/s"
    for key in ("value", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "cpu_baseline", "e2e"):
        assert key in d
    ncpu = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["cores"] == ncpu and d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["value"] > 0
    for dirpath, _, files in os.walk(os.path.join(root, "rendertoy_b200")):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f"{f} imports the oracle"
