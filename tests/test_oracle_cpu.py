"""CPU checks of the oracle itself (no GPU): known-answer values, structural invariants, BVH vs brute force."""
import numpy as np

from rendertoy_b200 import scenes


def _bits(a):
    return np.asarray(a, np.float32).ravel().view(np.uint32)


def _globals(hm, t=0.5, eye_z=1.0, aspect=1920 / 1080):
    W = hm.matmul(hm.scale(1.0), hm.rotate(t, (0, 1, 0)))
    V = hm.look_at((0, 0.3, eye_z), (0, 0, 0), (0, 1, 0))
    P = hm.perspective(aspect_ratio=aspect)
    return np.concatenate([W.ravel(), V.ravel(), P.ravel()]).astype(np.float32)


def test_camera_known_answers(oracle):
    """host matrices with NumPy-1 casting, bit for bit."""
    hm = oracle.host_math
    V = hm.look_at((0, 0.3, 1.0), (0, 0, 0), (0, 1, 0))
    assert _bits(V[1, 1]) == 0x3F75341A and _bits(V[1, 2]) == 0xBE931F43 and _bits(V[2, 2]) == 0xBF75341A
    assert _bits(V[3, 2]) == 0x3F85A2CC and _bits(V[0, 0]) == 0xBF800000
    V6 = hm.look_at((0, 0.3, 2), (0, 0, 0), (0, 1, 0))
    assert _bits(V6[1, 1]) == 0x3F7D2AEF and _bits(V6[1, 2]) == 0xBE17E690 and _bits(V6[3, 2]) == 0x40016E97
    P = hm.perspective(aspect_ratio=1920 / 1080)
    assert [int(x) for x in _bits([P[0, 0], P[1, 1], P[2, 2], P[3, 2]])] == [0x3FADD2C7, 0x401A8278, 0x3F800347, 0xBC23DB3C]
    assert _bits(hm.perspective(aspect_ratio=1.0)[0, 0]) == 0x401A8278
    W = hm.rotate(0.5, (0, 1, 0))
    assert [int(x) for x in _bits([W[0, 0], W[0, 2], W[2, 0]])] == [0x3F60A940, 0xBEF57744, 0x3EF57744]


def test_vertex_known_answer(oracle):
    """SURVEY.md appendix D end-to-end vertex KAT: strict no-FMA mul chain + dehomogenize + viewport."""
    clip, scr = oracle.vertex_kat((0.1, 0.2, -0.3), _globals(oracle.host_math), 1920, 1080)
    assert [int(x) for x in _bits(clip)] == [0x3D9BF054, 0x3F2DAA19, 0x3FA32C04, 0x3FA46F84]
    assert [int(x) for x in _bits(scr)] == [0x447E3994, 0x437ED8B0, 0x3F7E085D, 0x3FA46F84]


def _soup(tris):
    rows = np.zeros((len(tris) * 3, 20), np.float32)
    rows[:, 0:3] = np.asarray(tris, np.float32).reshape(-1, 3)
    rows[:, 4:7] = (0.0, 0.0, 1.0)
    return rows


def test_depth_is_min_and_ties_go_to_lowest_primitive(oracle):
    g = _globals(oracle.host_math, t=0.0, aspect=1.0)
    tri = [(-0.2, -0.2, 0.0), (0.2, -0.2, 0.0), (0.0, 0.25, 0.0)]
    behind = [(-0.3, -0.3, -0.2), (0.3, -0.3, -0.2), (0.0, 0.35, -0.2)]   # farther from the eye at z=+1
    r = oracle.draw_triangles(8, 128, 128, _soup([behind, tri, tri]), g)
    covered = r.winner != oracle.NO_WINNER
    assert covered.any()
    inner = r.winner == 2          # the first copy of the front triangle (primitive id 2*1+0) beats its twin (id 4)
    assert inner.any() and not (r.winner == 4).any()
    assert r.stats["tie_pixels"] == int(inner.sum())
    assert (r.depth[inner] < r.depth[r.winner == 0].min()).all() if (r.winner == 0).any() else True
    assert (r.depth[~covered] == 0x3F800000).all() and (r.bgra[~covered] == 0).all()


def test_near_clip_cases_and_large_triangle_drop(oracle):
    hm = oracle.host_math
    g = _globals(hm, t=0.0, aspect=1.0)
    rng = np.random.default_rng(11)
    # triangles straddling the near plane in every pattern: eye at z=1 looking down -z, near plane at z ~ 0.99
    tris = []
    for code in range(8):
        zs = [1.2 if code & (1 << k) else 0.5 for k in range(3)]
        xy = rng.uniform(-0.05, 0.05, (3, 2))
        tris.append([(xy[k, 0], xy[k, 1] + 0.25, zs[k]) for k in range(3)])
    r = oracle.draw_triangles(8, 256, 256, _soup(tris), g)
    assert r.stats["triangles_in"] == 8
    assert r.stats["primitives"] == 1 + 3 * 2 + 3 * 1          # code 0: 1, one vertex behind: 2 each, two behind: 1 each, 7: none
    big = [(-5.0, -5.0, 0.0), (5.0, -5.0, 0.0), (0.0, 5.0, 0.0)]   # covers the whole 256x256 target: bbox >= 64*64 -> dropped
    r2 = oracle.draw_triangles(8, 256, 256, _soup([big]), g)
    assert r2.stats["dropped_large"] >= 1 and r2.stats["fragments"] == 0 and (r2.winner == oracle.NO_WINNER).all()


def test_indexed_equals_soup(oracle):
    rows = scenes.dragon(600)
    g = _globals(oracle.host_math)
    idx = np.arange(rows.shape[0], dtype=np.int32)
    rng = np.random.default_rng(2)
    perm = rng.permutation(rows.shape[0] // 3)
    a = oracle.draw_triangles(8, 200, 150, rows, g)
    shuffled = idx.reshape(-1, 3)[perm].ravel()
    b = oracle.draw_triangles(8, 200, 150, rows, g, indices=shuffled)
    assert np.array_equal(a.depth, b.depth)
    same_colour = (a.bgra == b.bgra).all(axis=-1)
    assert same_colour[a.tie == 0].all()          # only depth ties may pick another primitive when ids are permuted


def test_texture_shader_runs(oracle):
    rows = scenes.dragon(600)
    tex = np.ones((7, 5, 4), np.float32)
    tex[..., 0:3] = np.random.default_rng(1).random((7, 5, 3), dtype=np.float32)
    r = oracle.draw_triangles(9, 160, 120, rows, _globals(oracle.host_math), texture=tex)
    assert r.stats["pixels_written"] > 0 and (r.bgra[r.winner != oracle.NO_WINNER][:, 3] == 255).all()


def test_cpu_bvh_matches_bruteforce(oracle):
    rows = scenes.dragon(3000)
    hm = oracle.host_math
    cam = hm.camera_frame(hm.look_at((0, 0.3, 2), (0, 0, 0), (0, 1, 0)), hm.perspective(aspect_ratio=4 / 3), hm.rotate(0.5, (0, 1, 0)))
    rays = oracle.primary_rays(cam, 160, 120)
    bt, bi, bu, bv = oracle.raycast_brute(rows, rays)
    h = oracle.bvh_build(rows)
    t, i, u, v = oracle.bvh_raycast(h, rays)
    oracle.bvh_free(h)
    assert (bi != 0xFFFFFFFF).sum() > 1000
    assert np.array_equal(i, bi) and np.array_equal(t.view(np.uint32), bt.view(np.uint32))
    assert np.array_equal(u.view(np.uint32), bu.view(np.uint32)) and np.array_equal(v.view(np.uint32), bv.view(np.uint32))


def test_raster_and_raycast_agree_on_visibility(oracle):
    """The one reference-anchored check of the ray caster: same camera, same winners."""
    rows = scenes.dragon(3000)
    hm = oracle.host_math
    w, h = 200, 150
    W, V, P = hm.rotate(0.5, (0, 1, 0)), hm.look_at((0, 0.3, 2), (0, 0, 0), (0, 1, 0)), hm.perspective(aspect_ratio=w / h)
    r = oracle.draw_triangles(8, w, h, rows, np.concatenate([W.ravel(), V.ravel(), P.ravel()]))
    t, ids, u, v = oracle.raycast_brute(rows, oracle.primary_rays(hm.camera_frame(V, P, W), w, h))
    ras = np.where(r.winner == oracle.NO_WINNER, 0xFFFFFFFF, r.winner // 2).ravel()
    both = (ras != 0xFFFFFFFF) & (ids != 0xFFFFFFFF)
    assert both.sum() > 1000 and (ras == ids)[both].mean() > 0.97
    shaded = oracle.shade_hits(8, rows, ids, u, v).reshape(h, w, 4).astype(int)
    same = (ras == ids) & both
    assert np.abs(shaded - r.bgra.astype(int)).reshape(-1, 4)[same].max() <= 1     # RGB within 1/255 where the winner agrees
