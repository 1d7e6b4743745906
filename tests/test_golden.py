"""Golden vectors produced by the UNMODIFIED reference (its Raster class + its own OpenCL C kernel text, executed on
the CPU through oracle/clshim; generator: oracle/clshim/make_golden.py).  CPU: the oracle must reproduce them.
GPU: the product must reproduce them, through the public API and the C ABI."""
import glob
import os

import numpy as np
import pytest

from rendertoy_b200 import lessons

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raster_*.npz")))


def _load(path):
    z = np.load(path)
    n = int(z["n_draws"])
    rows = [z[f"rows{i}"] for i in range(n)]
    idx = [z[f"indices{i}"] if z[f"indices{i}"].size else None for i in range(n)]
    tex = z["texture"] if z["texture"].size else None
    return z, rows, idx, tex


def test_fixtures_present():
    assert len(GOLDEN) >= 6, "tests/golden/raster_*.npz missing: run oracle/clshim/make_golden.py"


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_oracle_reproduces_reference_run(oracle, path):
    z, rows, idx, tex = _load(path)
    depth = bgra = None
    tie = np.zeros(z["depth"].shape, np.uint8)
    for r, ix in zip(rows, idx):
        draw = oracle.draw_points if int(z["points"]) else oracle.draw_triangles
        res = draw(int(z["lesson"]), int(z["width"]), int(z["height"]), r, z["globals"], indices=ix, texture=tex, depth=depth, bgra=bgra)
        depth, bgra = res.depth, res.bgra
        tie |= res.tie
    assert np.array_equal(depth, z["depth"]), "depth bits differ from the reference run"
    diff = (bgra != z["bgra"]).any(axis=-1)
    assert not (diff & (tie == 0)).any(), "colour differs from the reference run outside depth ties"


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[7:-4] for p in GOLDEN])
def test_product_reproduces_reference_run(ren, path):
    z, rows, idx, tex = _load(path)
    w, h, lesson = int(z["width"]), int(z["height"]), int(z["lesson"])
    pres = ren.create_presenter(w, h)
    if lesson == 8:
        raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    else:
        tex8 = np.rint(tex[:, :, 0:3] * 255.0).astype(np.uint8)     # the fixture stores image/255.0 as the tutorial does
        raster, g, _, _ = lessons.build_lesson09(ren, pres.get_render_target(), tex8)
    gl = z["globals"].reshape(3, 16)
    with ren.mapped(g) as m:
        m["World"], m["View"], m["Proj"] = (ren.make_float4x4(np.ascontiguousarray(x)) for x in gl)
    ren.clear(raster.get_render_target())
    ren.clear(raster.get_depth_buffer(), 1.0)
    for r, ix in zip(rows, idx):
        vb = ren.create_buffer(r.shape[0], ren.MeshVertex)
        with ren.mapped(vb) as m:
            m.view(np.float32).reshape(r.shape)[:] = r
        ib = None if ix is None else ren.create_buffer_from(np.ascontiguousarray(ix, np.int32))
        if int(z["points"]):
            raster.draw_points(vb, ib)
        else:
            raster.draw_triangles(vb, ib)
    depth = raster.get_depth_buffer().get().reshape(h, w)
    assert np.array_equal(depth, z["depth"]), "depth bits differ from the reference run"
    diff = (raster.get_render_target().get() != z["bgra"]).any(axis=-1)
    assert not (diff & (z["tie"] == 0)).any(), "colour differs from the reference run outside depth ties"


# ---- ray casting: no reference implementation exists (rendering/_raycaster.py:35-36 is `pass`), so the fixture comes from a
# second, independent restatement of the definition in numpy float32 (tests/golden/make_raycast_golden.py)

RAYCAST_GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "raycast_numpy_dragon300.npz")


def _same_hits(t, ids, u, v, z, label):
    assert np.array_equal(ids, z["ids"]), f"{label}: triangle ids differ on {int((ids != z['ids']).sum())} rays"
    for name, a in (("t", t), ("u", u), ("v", v)):
        assert np.array_equal(np.asarray(a, np.float32).view(np.uint32), z[name].view(np.uint32)), f"{label}: {name} bits differ"


def test_oracle_reproduces_numpy_raycast():
    import oracle
    z = np.load(RAYCAST_GOLDEN)
    w, h = int(z["width"]), int(z["height"])
    rays = oracle.primary_rays(z["camera"], w, h)
    t, ids, u, v = oracle.raycast_brute(z["rows"], rays)
    _same_hits(t, ids, u, v, z, "brute force")
    bvh = oracle.bvh_build(z["rows"])
    tb, ib, ub, vb = oracle.bvh_raycast(bvh, rays)
    oracle.bvh_free(bvh)
    _same_hits(tb, ib, ub, vb, z, "cpu bvh")
    assert np.array_equal(oracle.shade_hits(8, z["rows"], ids, u, v).reshape(h, w, 4), z["bgra"]), "Lambert BGRA8 differs"


@pytest.mark.gpu
@pytest.mark.parametrize("builder,view_nodes", [("ploc", True), ("lbvh", True), ("ploc", False)])
def test_cuda_reproduces_numpy_raycast(ren, builder, view_nodes):
    import torch
    from rendering._raycaster import Raycaster
    z = np.load(RAYCAST_GOLDEN)
    w, h = int(z["width"]), int(z["height"])
    rows = z["rows"]
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    rc = Raycaster([ren.Mesh(vb, None)], builder=builder)
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((w * h, 4), dtype=torch.float32, device="cuda")
    rc.render(target, z["camera"], hits=hits, view_nodes=view_nodes)
    got = hits.cpu().numpy()
    _same_hits(got[:, 0], got[:, 1].view(np.uint32), got[:, 2], got[:, 3], z, f"cuda {builder} view_nodes={view_nodes}")
    assert np.array_equal(target.get(), z["bgra"]), "Lambert BGRA8 differs"
