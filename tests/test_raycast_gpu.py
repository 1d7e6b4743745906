"""GPU parity: rendering._raycaster.Raycaster (LBVH build + traversal + shade, native sm_100a) vs the oracle's
brute-force float32 Moller-Trumbore definition (small scenes) and its CPU BVH (full-size scenes)."""
import os

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


def _validate_tree(rc, label):
    """Download the node array and check it is a proper binary tree over all leaves, rooted at 0, with child boxes that
    contain everything below them.  Returns the height.  (Done before any traversal: a malformed tree could hang one.)"""
    n = rc.n_triangles
    nodes = rc.nodes.cpu().numpy().view(np.float32).reshape(-1, 16)[:max(n - 1, 1)]
    tris = rc.tris.cpu().numpy().view(np.float32).reshape(-1, 12)[:n]
    child = nodes.view(np.int32)[:, 12:14]
    assert ((child >= -n) & (child < max(n - 1, 1))).all(), f"{label}: child reference out of range"
    # leaf boxes straight from the triangles (unpadded): every ancestor box must contain them
    v0, v1, v2 = tris[:, 0:3], tris[:, 0:3] + tris[:, 4:7], tris[:, 0:3] + tris[:, 8:11]
    seen_leaf = np.zeros(n, dtype=np.int32)
    seen_node = np.zeros(max(n - 1, 1), dtype=np.int32)
    height = 0
    stack = [(0, 1, None)]   # node, depth, box it must fit in: (lo, hi)
    while stack:
        node, depth, bound = stack.pop()
        seen_node[node] += 1
        assert seen_node[node] == 1, f"{label}: node {node} reached twice"
        height = max(height, depth)
        r = nodes[node]
        boxes = [(np.array([r[0], r[2], r[8]]), np.array([r[1], r[3], r[9]])), (np.array([r[4], r[6], r[10]]), np.array([r[5], r[7], r[11]]))]
        for k in range(2):
            lo, hi = boxes[k]
            c = int(child[node, k])
            if n == 1 and k == 1:
                continue                      # the single-triangle root's empty second child
            if bound is not None and np.isfinite(lo).all() and np.isfinite(hi).all():
                assert (lo >= bound[0]).all() and (hi <= bound[1]).all(), f"{label}: child box of node {node} sticks out of its parent's"
            if c < 0:
                s = ~c
                seen_leaf[s] += 1
                pts = np.stack([v0[s], v1[s], v2[s]])
                if np.isfinite(pts).all():
                    assert (pts.min(0) >= lo).all() and (pts.max(0) <= hi).all(), f"{label}: leaf {s} outside its box"
            else:
                stack.append((c, depth + 1, (lo, hi)))
    assert (seen_leaf == 1).all(), f"{label}: {int((seen_leaf != 1).sum())} leaves not reached exactly once"
    if n > 1:
        assert (seen_node == 1).all(), f"{label}: unreachable inner nodes"
    return height


def test_ploc_builder_trees_and_hits(ren, oracle):
    """rt_bvh_build(RT_BVH_PLOC): structurally valid trees (checked on the host before anything traverses them) for regular,
    tiny, duplicate-heavy, skewed and non-finite inputs; the same hit records as the Karras tree and as brute force."""
    from rendering._raycaster import Raycaster
    rng = np.random.default_rng(5)
    base = scenes.dragon(3_000)
    scenes_ = {"dragon3k": base}
    for k in (1, 2, 3, 4, 5, 33):
        scenes_[f"first {k} triangles"] = base[:3 * k].copy()
    dup = np.tile(base[:3], (700, 1))                           # 700 copies of one triangle: all Morton codes equal
    scenes_["700 identical triangles"] = dup
    chain = base[:3 * 400].copy()                                # geometrically growing triangles along a line: nearest
    for t in range(400):                                         # neighbour is always the smaller one -- worst case for mutual pairs
        chain[3 * t:3 * t + 3, 0:3] = base[0:3, 0:3] * (1.03 ** t) + np.array([1.03 ** t, 0, 0])
    scenes_["growing chain"] = chain
    bad = base.copy()
    bad[rng.integers(0, bad.shape[0], 12), rng.integers(0, 3, 12)] = np.nan
    bad[rng.integers(0, bad.shape[0], 6), rng.integers(0, 3, 6)] = np.inf
    scenes_["non-finite vertices"] = bad
    w, h = 128, 72
    for label, rows in scenes_.items():
        vb = _mesh_buffer(ren, np.ascontiguousarray(rows))
        rc_p = Raycaster([ren.Mesh(vb, None)], builder="ploc")
        rc_l = Raycaster([ren.Mesh(vb, None)], builder="lbvh")
        hp, hl = _validate_tree(rc_p, label + " (ploc)"), _validate_tree(rc_l, label + " (lbvh)")
        assert hp <= 64 and hl <= 64, f"{label}: tree heights {hp} / {hl} exceed the traversal stack"
        finite = rows[np.isfinite(rows[:, :3]).all(axis=1), :3]
        centre, ext = (finite.min(0) + finite.max(0)) / 2, float((finite.max(0) - finite.min(0)).max())
        eye = centre + np.array([0.3, 0.4, -2.0]) * max(ext, 1e-3)
        fwd = (centre - eye) / np.linalg.norm(centre - eye)
        right = np.cross([0, 1, 0], fwd); right /= np.linalg.norm(right)
        up = np.cross(fwd, right)
        cam = np.concatenate([eye, right * 0.5 * w / h, up * 0.5, fwd]).astype(np.float32)
        a, b = _hits(rc_p, cam, w, h), _hits(rc_l, cam, w, h)
        assert np.array_equal(a, b), f"{label}: PLOC and LBVH trees give different hits on {int((a != b).any(axis=1).sum())} pixels"
        ref = oracle.raycast_brute(np.ascontiguousarray(rows), oracle.primary_rays(cam, w, h))
        _check(a.view(np.float32), ref, label)
        print(f"{label}: n={rows.shape[0] // 3} heights ploc {hp} lbvh {hl}, hit pixels {int((a[:, 1] != 0xFFFFFFFF).sum())}")


def test_render_content_rect_holds_every_hit(ren):
    """render() returns the rect outside of which the frame is the clear colour: the contract of the sparse gather."""
    from rendering._raycaster import Raycaster
    rows = scenes.dragon(5_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    for lesson, w, h, t in [(6, 640, 360, 0.3), (8, 333, 211, 2.1), (6, 1920, 1080, 4.0), (8, 96, 64, 5.5)]:
        cam = _camera(ren, lesson, t, w, h)
        target = ren.create_image2d(w, h, ren._core.RGBA)
        target.buffer.tensor().fill_(0x5A)                           # stale pixels everywhere
        x0, y0, x1, y1 = rc.render(target, cam)
        img = target.get().view(np.uint32).reshape(h, w)
        assert (img != 0).any()
        outside = np.ones((h, w), bool)
        outside[y0:y1 + 1, x0:x1 + 1] = False
        assert not img[outside].any(), "non-clear pixel outside the returned content rect"
        sub = rc.render(target, cam, rect=(8, 4, w - 16, h - 8))     # partial rect: content is clipped to it
        assert sub[0] >= 8 and sub[1] >= 4 and sub[2] <= w - 9 and sub[3] <= h - 5
    # camera inside the scene: no bound, the content rect is the whole frame
    cam = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)
    target = ren.create_image2d(64, 48, ren._core.RGBA)
    assert rc.render(target, cam) == (0, 0, 63, 47)


def test_frame_store_sparse_push(ren):
    """FrameStore.push (rt_copy_rect): an orbit of locally rendered frames cycling through two slots of a cleared store;
    after every push the slot equals the frame bit for bit although only the cover rect travelled."""
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import parallel
    w, h = 1280, 720
    rows = scenes.dragon(20_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    store = parallel.FrameStore(2, w, h)
    assert store.ok
    try:
        assert not store.frames().any(), "the store must start cleared"
        local = ren.create_image2d(w, h, ren._core.RGBA)
        side = torch.cuda.Stream()
        moved = 0
        for k in range(12):
            cam = _camera(ren, 6 if k % 3 else 8, 0.45 * k, w, h)
            content = rc.render(local, cam)
            side.wait_stream(torch.cuda.current_stream())
            moved += store.push(k % 2, local.ptr, content, side.cuda_stream)
            torch.cuda.current_stream().wait_stream(side)
            frame = local.buffer.tensor().view(torch.int32).view(h, w)
            assert torch.equal(store.frames()[k % 2], frame), f"slot differs from frame {k}"
            assert frame.any()
        assert 0 < moved < 0.7 * 12 * w * h * 4
    finally:
        store.close()


def test_view_refit_keeps_hits(ren):
    """rt_raycast_set_view_refit(k): tightening passes over the screen-space nodes must not change a single hit record."""
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import _native
    rows = scenes.dragon(30_000)
    w, h = 1280, 720
    try:
        for builder in ("ploc", "lbvh"):
            rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)], builder=builder)
            for lesson, t in [(6, 0.5), (8, 2.2), (6, 4.1)]:
                cam = _camera(ren, lesson, t, w, h)
                ref = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
                _native.call("rt_raycast_set_view_refit", 0)
                st0 = torch.zeros(3, dtype=torch.int64, device="cuda")
                rc.render(None, cam, hits=ref, frame_size=(w, h), stats=st0)
                for passes in (1, 2, 4, 16):
                    _native.call("rt_raycast_set_view_refit", passes)
                    got = torch.empty_like(ref)
                    st = torch.zeros(3, dtype=torch.int64, device="cuda")
                    rc.render(None, cam, hits=got, frame_size=(w, h), stats=st)
                    assert torch.equal(got.view(torch.int32), ref.view(torch.int32)), f"{builder}, {passes} passes: hits changed"
                    # (not strictly monotone: tighter depth bounds can swap the visiting order of two children)
                    assert int(st[0]) <= 1.02 * int(st0[0]), "tightening should not add node visits"
                    print(builder, lesson, passes, "node visits", int(st0[0]), "->", int(st[0]))
    finally:
        _native.call("rt_raycast_set_view_refit", 0)


def _texture_pool(ren, w, h, seed=11):
    rgb = np.random.default_rng(seed).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    mem, desc = ren.create_texture2D(w, h)
    with ren.mapped(mem) as m:
        m = m.view(np.float32).ravel().reshape(h, w, 4)
        m[:, :, 0:3] = rgb / 255.0
        m[:, :, 3] = 1.0
        texf = np.array(m)
    return desc, texf


@pytest.mark.parametrize("w,h,lesson,t,shader", [
    (3840, 2160, 6, 0.5, 8),      # configs[3] as quoted: dragon100k at 4K, lesson06 camera
    (3840, 2160, 8, 2.2, 8),      # the frame-filling lesson08 camera at 4K
    (1920, 1080, 8, 1.3, 9),      # configs[2]: texture-mapped ray cast at 1080p, marble2.jpg-sized (500x500) texture
])
def test_quoted_configs_match_cpu_bvh(ren, oracle, w, h, lesson, t, shader):
    """Parity AT the sizes BASELINE.json quotes: every hit record of the frame against the oracle's CPU BVH
    (>= 99.99 % of the ids, t/u/v bit-identical where the id agrees -- the oracle's BVH walk and the brute-force definition
    differ only at exact t ties between adjacent triangles) and the shaded BGRA8 frame wherever the hit agrees."""
    from rendering._raycaster import Raycaster
    rows = scenes.dragon(100_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    cam = _camera(ren, lesson, t, w, h)
    desc = texf = None
    if shader == 9:
        desc, texf = _texture_pool(ren, 500, 500)
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(target, cam, shader=shader, texture_descriptor=desc, hits=hits)
    bvh = oracle.bvh_build(rows)
    ref = oracle.bvh_raycast(bvh, oracle.primary_rays(cam, w, h))
    oracle.bvh_free(bvh)
    got = hits.cpu().numpy()
    agree = _check(got, ref, f"cpu-bvh dragon100k {w}x{h} lesson{lesson:02d} camera, shader {shader}", min_agree=0.9999)
    assert (ref[1] != 0xFFFFFFFF).mean() > 0.05
    shaded = oracle.shade_hits(shader, rows, ref[1], ref[2], ref[3], texture=texf).reshape(h, w, 4)
    diff = (target.get() != shaded).any(axis=-1).reshape(-1)
    assert not diff[agree].any(), "shaded colour differs where the hit agrees"


@pytest.mark.parametrize("w,h,n_tris,world,lesson,view_nodes", [
    (3840, 2160, 100_000, 8, 6, None),    # configs[3]: one 4K frame over 8 ranks
    (3840, 2160, 100_000, 3, 8, None),
    (1000, 600, 20_000, 2, 8, False),     # the per-lane 3-D walk; height not a multiple of the band
])
def test_stripe_partition_equals_the_whole_frame(ren, w, h, n_tris, world, lesson, view_nodes):
    """configs[3]'s image-space partition: rank r traces the row stripes (parallel.BAND, world, r) of ONE frame with one
    launch; together the ranks must produce the unpartitioned frame byte for byte (hits and BGRA8), and a rank must not touch
    rows it does not own."""
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import parallel
    rows = scenes.dragon(n_tris)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    cam = _camera(ren, lesson, 0.9, w, h)
    whole = ren.create_image2d(w, h, ren._core.RGBA)
    hits_whole = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(whole, cam, hits=hits_whole, view_nodes=view_nodes)
    parts = ren.create_image2d(w, h, ren._core.RGBA)
    parts.buffer.tensor().fill_(0xAB)
    hits_parts = torch.full((w * h, 4), float("nan"), dtype=torch.float32, device="cuda")
    yy = torch.arange(h, device="cuda")
    for rank in range(world):
        before_c = parts.buffer.tensor().clone().view(h, w * 4)
        before_h = hits_parts.clone().view(torch.int32).view(h, w * 4)
        content = rc.render(parts, cam, hits=hits_parts, view_nodes=view_nodes, stripes=(parallel.BAND, world, rank))
        foreign = (yy // parallel.BAND) % world != rank
        assert torch.equal(parts.buffer.tensor().view(h, w * 4)[foreign], before_c[foreign]), f"rank {rank} wrote pixels of foreign stripes"
        assert torch.equal(hits_parts.view(torch.int32).view(h, w * 4)[foreign], before_h[foreign]), f"rank {rank} wrote hits of foreign stripes"
        assert content[0] >= 0 and content[2] < w
    assert torch.equal(parts.buffer.tensor(), whole.buffer.tensor()), "striped frame differs from the whole frame"
    assert torch.equal(hits_parts.view(torch.int32), hits_whole.view(torch.int32)), "striped hit records differ"
    assert (whole.get()[:, :, 3] != 0).any()


@pytest.mark.parametrize("w,h,world", [(3840, 2160, 8), (1920, 1080, 3), (500, 333, 2)])
def test_frame_store_stripe_push(ren, w, h, world):
    """FrameStore.push_stripes (rt_copy_stripes: one 3-D copy-engine transfer for a rank's whole stripes + 2-D ones for cut
    stripes): every "rank" pushes its stripes of an orbit of locally rendered frames into the same two slots; after all ranks
    pushed, the slot equals the frame bit for bit although only the cover rect travelled; a rank's push must not write foreign rows."""
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import parallel
    rows = scenes.dragon(20_000)
    rc = Raycaster([ren.Mesh(_mesh_buffer(ren, rows), None)])
    store = parallel.FrameStore(2, w, h)       # push_stripes keeps the slot's previous content rect per (slot, stripes)
    assert store.ok
    try:
        local = ren.create_image2d(w, h, ren._core.RGBA)
        side = torch.cuda.Stream()
        yy = torch.arange(h, device="cuda")
        for k in range(6):
            cam = _camera(ren, 6 if k % 3 else 8, 0.45 * k, w, h)
            content = rc.render(local, cam)
            frame = local.buffer.tensor().view(torch.int32).view(h, w)
            side.wait_stream(torch.cuda.current_stream())
            for rank in range(world):
                before = store.frames()[k % 2].clone()
                store.push_stripes(k % 2, local.ptr, content, (parallel.BAND, world, rank), side.cuda_stream)
                side.synchronize()
                foreign = (yy // parallel.BAND) % world != rank
                assert torch.equal(store.frames()[k % 2][foreign], before[foreign]), f"rank {rank} wrote foreign rows"
            assert torch.equal(store.frames()[k % 2], frame), f"slot differs from frame {k}"
            assert frame.any()
    finally:
        store.close()
