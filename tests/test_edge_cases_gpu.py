"""Edge cases on the GPU, each against the oracle: empty and ragged inputs, degenerate / NaN / far-off-screen geometry,
single primitives, odd viewports, off-centre depth clears."""
import numpy as np
import pytest
import torch

from rendertoy_b200 import lessons, scenes

pytestmark = pytest.mark.gpu


def _upload(ren, rows):
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    if rows.shape[0]:
        with ren.mapped(vb) as m:
            m.view(np.float32).reshape(rows.shape)[:] = rows
    return vb


def _soup(tris, normal=(0.0, 0.0, 1.0)):
    rows = np.zeros((len(tris) * 3, 20), np.float32)
    rows[:, 0:3] = np.asarray(tris, np.float32).reshape(-1, 3)
    rows[:, 4:7] = normal
    return rows


def _frame(ren, oracle, rows, w, h, t=0.0, depth_clear=1.0, lesson=8):
    pres = ren.create_presenter(w, h)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, lesson, t, w, h))
    lessons.render_frame(ren, raster, _upload(ren, rows), depth_clear=depth_clear)
    d0 = np.full((h, w), np.float32(depth_clear).view(np.uint32), np.uint32)
    res = oracle.draw_triangles(8, w, h, rows, lessons.globals_as_floats(g), depth=d0)
    depth = raster.get_depth_buffer().get().reshape(h, w)
    assert np.array_equal(depth, res.depth)
    diff = (raster.get_render_target().get() != res.bgra).any(axis=-1)
    assert not (diff & (res.tie == 0)).any()
    return res


def test_empty_and_ragged_vertex_buffers(ren, oracle):
    res = _frame(ren, oracle, np.zeros((0, 20), np.float32), 64, 48)
    assert res.stats["fragments"] == 0
    rows = scenes.dragon(300)
    for cut in (1, 2):                      # vertex count not a multiple of 3: the tail is ignored (:419 floor division)
        r = _frame(ren, oracle, rows[:rows.shape[0] - cut], 96, 64, t=0.5)
        assert r.stats["triangles_in"] == rows.shape[0] // 3 - 1
    one = _frame(ren, oracle, _soup([[(-0.2, -0.2, 0), (0.2, -0.2, 0), (0, 0.2, 0)]]), 33, 17)
    assert one.stats["pixels_written"] > 0


def test_degenerate_nan_and_offscreen_geometry(ren, oracle):
    rng = np.random.default_rng(9)
    tris = []
    tris += [[(0.1, 0.1, 0.0)] * 3]                                                   # a point
    tris += [[(-0.3, 0.0, 0.0), (0.0, 0.0, 0.0), (0.3, 0.0, 0.0)]]                    # collinear
    tris += [[(np.nan, 0.0, 0.0), (0.2, 0.1, 0.0), (0.0, 0.3, 0.0)]]                  # NaN vertex
    tris += [[(np.inf, 0.0, 0.0), (0.2, 0.1, 0.0), (0.0, 0.3, 0.0)]]                  # inf vertex
    tris += [[(50.0, 50.0, 0.0), (51.0, 50.0, 0.0), (50.0, 51.0, 0.0)]]               # far off-screen (both extents negative)
    tris += [[(-50.0, -50.0, 0.0), (-51.0, -50.0, 0.0), (-50.0, -51.0, 0.0)]]
    tris += [[(-0.4, -0.3, 0.2), (0.4, -0.3, 0.2), (0.0, 0.4, 0.2)]]                  # an ordinary one behind / in front of others
    tris += [[(-1e-4, -1e-4, 0.1), (1e-4, -1e-4, 0.1), (0.0, 1e-4, 0.1)]]             # sub-pixel
    tris += [rng.uniform(-0.5, 0.5, (3, 3)).tolist() for _ in range(200)]             # large overlapping random triangles
    tris += [[(0.0, 0.0, 1.5), (0.1, 0.0, 0.95), (0.0, 0.1, 0.5)]]                    # crosses the near plane and the eye plane
    rows = _soup(tris)
    for (w, h) in [(200, 150), (64, 64), (7, 5)]:
        _frame(ren, oracle, rows, w, h, t=0.3)
    _frame(ren, oracle, rows, 200, 150, t=0.3, depth_clear=0.97)                        # depth cleared below some fragments


def test_raycast_degenerate_inputs(ren, oracle):
    from rendering._raycaster import Raycaster, Ray
    tris = [[(0.1, 0.1, 0.0)] * 3, [(-0.3, 0.0, 0.0), (0.0, 0.0, 0.0), (0.3, 0.0, 0.0)],
            [(-0.4, -0.3, 0.2), (0.4, -0.3, 0.2), (0.0, 0.4, 0.2)], [(-0.4, -0.3, -0.2), (0.4, -0.3, -0.2), (0.0, 0.4, -0.2)]]
    rows = _soup(tris)
    rc = Raycaster([ren.Mesh(_upload(ren, rows), None)])
    n = 600
    rng = np.random.default_rng(2)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0:3] = rng.uniform(-1, 1, (n, 3)) + (0, 0, 2.0)
    rays[:, 4:7] = rng.uniform(-0.3, 0.3, (n, 3)) - rays[:, 0:3]
    rays[0, 4:7] = 0.0                       # zero direction
    rays[1, 4:7] = (0.0, 0.0, -1.0)          # axis aligned: two zero components
    rays[2, 4:7] = (np.nan, 0.0, -1.0)
    rays[3, 0:3], rays[3, 4:7] = (0.0, 0.0, 0.2), (0.0, 0.0, -1.0)   # origin ON a triangle's plane
    rb = ren.create_buffer(n, Ray)
    with ren.mapped(rb) as m:
        m.view(np.float32).reshape(n, 8)[:] = rays
    out = rc.ray_cast(rb).get()
    t, ids, u, v = oracle.raycast_brute(rows, rays)
    assert np.array_equal(out["index"], np.where(ids == 0xFFFFFFFF, -1, ids.astype(np.int64)).astype(np.int32))
    assert np.array_equal(out["t"].view(np.uint32), t.view(np.uint32))
    single = Raycaster([ren.Mesh(_upload(ren, _soup([tris[2]])), None)])        # one triangle: root with one real child
    out1 = single.ray_cast(rb).get()
    t1, i1, _, _ = oracle.raycast_brute(_soup([tris[2]]), rays)
    assert np.array_equal(out1["index"], np.where(i1 == 0xFFFFFFFF, -1, i1.astype(np.int64)).astype(np.int32))


def test_many_duplicate_centroids_in_bvh(ren, oracle):
    """Hundreds of triangles sharing one Morton code: the Karras index tie-break must still give a valid tree."""
    from rendering._raycaster import Raycaster
    base = np.array([(-0.2, -0.2, 0.0), (0.2, -0.2, 0.0), (0.0, 0.25, 0.0)], np.float32)
    tris = [base + (0, 0, 0.0001 * (k % 3)) for k in range(400)] + [base * 0.5 + (0.3, 0.3, -0.2)]
    rows = _soup(tris)
    rc = Raycaster([ren.Mesh(_upload(ren, rows), None)])
    w, h = 64, 48
    from rendering._raycaster import camera_frame
    world, view, proj = scenes.lesson_camera(ren, 6, 0.0, w, h)
    cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(None, cam, hits=hits, frame_size=(w, h))
    t, ids, u, v = oracle.raycast_brute(rows, oracle.primary_rays(cam, w, h))
    got = hits.cpu().numpy()
    assert np.array_equal(got[:, 1].view(np.uint32), ids) and (ids != 0xFFFFFFFF).any()
    assert np.array_equal(got[:, 0].view(np.uint32), t.view(np.uint32))
