"""GPU parity: rendering._raycaster.Raycaster (LBVH build + traversal + shade, native sm_100a) vs the oracle's
brute-force float32 Moller-Trumbore definition (small scenes) and its CPU BVH (full-size scenes)."""
import numpy as np
import pytest
import torch

from rendertoy_b200 import lessons, scenes

pytestmark = pytest.mark.gpu


def _mesh_buffer(ren, rows):
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    return vb


def _camera(ren, lesson, t, w, h):
    from rendering._raycaster import camera_frame
    world, view, proj = scenes.lesson_camera(ren, lesson, t, w, h)
    return camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))


def _check(hits, ref, label, min_agree=1.0):
    t, ids, u, v = ref
    got_t, got_id = hits[:, 0], hits[:, 1].view(np.uint32)
    agree = got_id == ids
    frac = agree.mean()
    same_t = (got_t.view(np.uint32) == t.view(np.uint32))[agree].mean()
    same_uv = ((hits[:, 2].view(np.uint32) == u.view(np.uint32)) & (hits[:, 3].view(np.uint32) == v.view(np.uint32)))[agree].mean()
    print(f"{label}: rays={len(ids)} hit={(ids != 0xFFFFFFFF).mean():.3f} id_agree={frac:.6f} t_bits={same_t:.6f} uv_bits={same_uv:.6f}")
    assert frac >= min_agree, f"{label}: triangle ids agree on {frac:.6f} < {min_agree}"
    assert same_t == 1.0 and same_uv == 1.0, f"{label}: t/u/v bits differ on agreeing rays"
    return agree


@pytest.mark.parametrize("n_tris,w,h,lesson", [(2_000, 96, 64, 6), (500, 64, 48, 8), (12, 40, 30, 6)])
def test_primary_rays_match_bruteforce(ren, oracle, n_tris, w, h, lesson):
    from rendering._raycaster import Raycaster
    rows = scenes.dragon(n_tris)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    cam = _camera(ren, lesson, 0.5, w, h)
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(target, cam, hits=hits)
    rays = oracle.primary_rays(cam, w, h)
    ref = oracle.raycast_brute(rows, rays)
    _check(hits.cpu().numpy(), ref, f"brute T={rows.shape[0] // 3} {w}x{h}")
    shaded = oracle.shade_hits(8, rows, ref[1], ref[2], ref[3]).reshape(h, w, 4)
    assert np.array_equal(target.get(), shaded), "Lambert BGRA8 differs from the oracle"


def test_textured_shading_matches_oracle(ren, oracle):
    """configs[2]: lesson09-style shading of ray-cast hits (texture over P.xy*2, L = 0.2 + max(0, N.l))."""
    from rendering._raycaster import Raycaster
    from rendertoy_b200._native import SHADER_LESSON09
    rows = scenes.dragon(3_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    w, h = 160, 120
    cam = _camera(ren, 8, 1.3, w, h)
    rgb = np.random.default_rng(3).integers(0, 256, size=(31, 45, 3), dtype=np.uint8)
    mem, desc = ren.create_texture2D(45, 31)
    with ren.mapped(mem) as m:
        m = m.view(np.float32).ravel().reshape(31, 45, 4)
        m[:, :, 0:3] = rgb / 255.0
        m[:, :, 3] = 1.0
        texf = np.array(m)
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(target, cam, shader=SHADER_LESSON09, texture_descriptor=desc, hits=hits)
    ref = oracle.raycast_brute(rows, oracle.primary_rays(cam, w, h))
    _check(hits.cpu().numpy(), ref, "textured")
    shaded = oracle.shade_hits(9, rows, ref[1], ref[2], ref[3], texture=texf).reshape(h, w, 4)
    assert (ref[1] != 0xFFFFFFFF).sum() > 2000
    assert np.array_equal(target.get(), shaded), "textured BGRA8 differs from the oracle"


def test_ray_buffer_matches_bruteforce(ren, oracle):
    from rendering._raycaster import Raycaster, Ray
    rows = scenes.dragon(1_500)
    rng = np.random.default_rng(5)
    n = 5000
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-1.5, 1.5, (n, 3))
    target = rows[rng.integers(0, rows.shape[0], n), 0:3] + rng.normal(0, 0.02, (n, 3))
    rays[:, 4:7] = target - rays[:, 0:3]
    rays[::7, 4] = 0.0   # axis-parallel components exercise the 0 * inf slab case
    rb = ren.create_buffer(n, Ray)
    with ren.mapped(rb) as m:
        m.view(np.float32).reshape(n, 8)[:] = rays
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    out = rc.ray_cast(rb).get()
    t, ids, u, v = oracle.raycast_brute(rows, rays)
    exp_index = np.where(ids == 0xFFFFFFFF, -1, ids.astype(np.int64)).astype(np.int32)
    assert np.array_equal(out["index"], exp_index)
    assert np.array_equal(out["t"].view(np.uint32), t.view(np.uint32))
    assert np.array_equal(out["mesh"], np.where(exp_index < 0, -1, 0))


def test_two_meshes_and_indices(ren, oracle):
    from rendering._raycaster import Raycaster
    a = scenes.dragon(600)
    grid = ren.manifold(6, 5)                                   # indexed int32 mesh, z = 0 plane
    with ren.mapped(grid.vertices) as m:
        g = m.view(np.float32).reshape(-1, 20)
        g[:, 0:2] -= 0.5
        g[:, 2] = -0.45
        g[:, 6] = 1.0
        grid_rows = g.copy()
    idx = grid.indices.get()
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, a), None), grid])
    w, h = 80, 60
    cam = _camera(ren, 6, 0.0, w, h)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(None, cam, hits=hits, frame_size=(w, h))
    # oracle on the flattened soup: mesh a then the grid's indexed triangles expanded
    flat = np.concatenate([a, grid_rows[idx]])
    ref = oracle.raycast_brute(flat, oracle.primary_rays(cam, w, h))
    _check(hits.cpu().numpy(), ref, "two meshes")
    assert (ref[1][ref[1] != 0xFFFFFFFF] >= a.shape[0] // 3).any(), "the grid mesh should be visible"


def test_full_size_matches_cpu_bvh_and_raster(ren, oracle):
    """dragon100k at 512x512 (BASELINE config 1) against the oracle's CPU BVH, then the raster winner cross-check
    (the one reference-anchored check the ray caster has: SURVEY.md section 8c)."""
    from rendering._raycaster import Raycaster
    rows = scenes.dragon(100_000)
    vb = _mesh_buffer(ren, rows)
    rc = Raycaster([ren.Mesh(vb, None)])
    w = h = 512
    cam = _camera(ren, 6, 0.5, w, h)
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(target, cam, hits=hits)
    rays = oracle.primary_rays(cam, w, h)
    bvh = oracle.bvh_build(rows)
    ref = oracle.bvh_raycast(bvh, rays)
    oracle.bvh_free(bvh)
    got = hits.cpu().numpy()
    agree = _check(got, ref, "cpu-bvh dragon100k 512x512", min_agree=0.9999)
    shaded = oracle.shade_hits(8, rows, ref[1], ref[2], ref[3]).reshape(h, w, 4)
    diff = (target.get() != shaded).any(axis=-1).reshape(-1)
    assert not diff[agree].any(), "shaded colour differs where the hit agrees"
    # tiles == full frame, byte for byte
    tiled = ren.create_image2d(w, h, ren._core.RGBA)
    for (x0, y0, tw, th) in [(0, 0, 200, 512), (200, 0, 312, 100), (200, 100, 312, 412)]:
        rc.render(tiled, cam, rect=(x0, y0, tw, th))
    assert np.array_equal(tiled.get(), target.get())
    # raster winner vs ray-cast hit under the same camera
    pres = ren.create_presenter(w, h)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 6, 0.5, w, h))
    lessons.render_frame(ren, raster, vb)
    res = oracle.draw_triangles(8, w, h, rows, lessons.globals_as_floats(g))
    ray_id = got[:, 1].view(np.uint32).reshape(h, w)
    ras_id = np.where(res.winner == 0xFFFFFFFF, 0xFFFFFFFF, res.winner // 2)
    both = (ray_id != 0xFFFFFFFF) & (ras_id != 0xFFFFFFFF)
    same = (ray_id == ras_id)[both].mean()
    cover = ((ray_id != 0xFFFFFFFF) == (ras_id != 0xFFFFFFFF)).mean()
    print(f"raster-vs-raycast: same triangle on {same:.4f} of commonly covered pixels, coverage agreement {cover:.4f}")
    assert same > 0.97 and cover > 0.99


def _hits(rc, cam, w, h, **kw):
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(None, cam, hits=hits, frame_size=(w, h), **kw)
    return hits.cpu().numpy().view(np.uint32)


def test_view_space_nodes_give_the_same_hits(ren, oracle):
    """render() projects the BVH into the camera's screen space once per frame and walks that (rectangle tests);
    the hit records must be bit-identical to the 3-D slab traversal for every camera -- also with the eye inside
    the scene's bounds (nodes straddling the eye plane), grazing views and a stretched camera basis."""
    from rendering._raycaster import Raycaster
    rows = scenes.dragon(20_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    w, h = 640, 360
    cams = [_camera(ren, 6, t, w, h) for t in (0.0, 0.7, 2.1)] + [_camera(ren, 8, 1.3, w, h)]
    inside = np.array(cams[0], dtype=np.float32).copy()
    inside[0:3] = (0.02, 0.01, 0.03)                        # eye in the middle of the mesh
    cams.append(inside)
    skew = np.array(cams[1], dtype=np.float32).copy()
    skew[3:6] *= 7.0                                        # very wide, anisotropic basis
    skew[6:9] *= 0.05
    cams.append(skew)
    for i, cam in enumerate(cams):
        a = _hits(rc, cam, w, h, view_nodes=True)
        b = _hits(rc, cam, w, h, view_nodes=False)
        assert np.array_equal(a, b), f"camera {i}: {int((a != b).any(axis=1).sum())} pixels differ"
    # brute force on a subset, so the two paths are not merely equal to each other
    ref = oracle.raycast_brute(rows, oracle.primary_rays(inside, 64, 36))
    _check(_hits(rc, inside, 64, 36, view_nodes=True).view(np.float32), ref, "eye inside the mesh, view nodes")


def test_view_space_nodes_degenerate_scenes(ren, oracle):
    from rendering._raycaster import Raycaster
    w, h = 96, 64
    cam = _camera(ren, 6, 0.4, w, h)
    one = scenes.dragon(12)[:3].copy()                      # a single triangle: root with an empty second child
    bad = scenes.dragon(300).copy()
    bad[7, 0] = np.nan                                      # non-finite vertices poison some boxes
    bad[40, 1] = np.inf
    bad[91, 2] = -np.inf
    for label, rows in (("single triangle", one), ("non-finite vertices", bad)):
        rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
        a = _hits(rc, cam, w, h, view_nodes=True)
        b = _hits(rc, cam, w, h, view_nodes=False)
        assert np.array_equal(a, b), label
        ref = oracle.raycast_brute(rows, oracle.primary_rays(cam, w, h))
        _check(a.view(np.float32), ref, label)


def test_view_space_nodes_random_cameras(ren, oracle):
    """Fuzz: 60 random camera frames (eye anywhere from inside the mesh to 40 extents away, arbitrary orientation and
    roll, field of view from 2 to 170 degrees, anisotropic and sheared bases, odd frame sizes and sub-rectangles):
    the screen-space packet traversal and the 3-D slab traversal must return bit-identical hit records."""
    from rendering._raycaster import Raycaster
    rng = np.random.default_rng(20240607)
    rows = scenes.dragon(6_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    checked = hit_pixels = 0
    for i in range(60):
        w, h = int(rng.integers(17, 200)), int(rng.integers(9, 120))
        eye = rng.normal(size=3)
        eye *= (rng.choice([0.02, 0.3, 0.8, 2.0, 12.0, 40.0]) / np.linalg.norm(eye))
        look = rng.normal(size=3) * 0.2 - eye if rng.random() < 0.8 else rng.normal(size=3)
        fwd = look / np.linalg.norm(look)
        up = rng.normal(size=3)
        right = np.cross(up, fwd); right /= np.linalg.norm(right)
        up = np.cross(fwd, right)
        tan_half = np.tan(np.radians(rng.choice([2.0, 20.0, 45.0, 90.0, 170.0])) / 2)
        U, V, W = right * tan_half * (w / h), up * tan_half, fwd.copy()
        if rng.random() < 0.3:      # shear / anisotropy: still a basis, no longer orthogonal
            U = U + 0.3 * V; W = W + 0.1 * U; V = V * rng.uniform(0.2, 3.0)
        cam = np.concatenate([eye, U, V, W]).astype(np.float32)
        rect = None
        if rng.random() < 0.4:
            x0, y0 = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 4))
            rect = (x0, y0, int(rng.integers(1, w - x0 + 1)), int(rng.integers(1, h - y0 + 1)))
        rw, rh = (rect[2], rect[3]) if rect else (w, h)
        out = []
        for vn in (True, False):
            hits = torch.empty((rw * rh, 4), dtype=torch.float32, device="cuda")
            rc.render(None, cam, rect=rect, hits=hits, frame_size=(w, h), view_nodes=vn, cull=bool(i % 2))
            out.append(hits.cpu().numpy().view(np.uint32))
        assert np.array_equal(out[0], out[1]), f"camera {i}: {int((out[0] != out[1]).any(axis=1).sum())} of {rw * rh} pixels differ"
        checked += rw * rh
        hit_pixels += int((out[0][:, 1] != 0xFFFFFFFF).sum())
    print(f"{checked} pixels over 60 cameras, {hit_pixels} hits")
    assert hit_pixels > 0.02 * checked, "the fuzz cameras should see the mesh often enough to mean something"
