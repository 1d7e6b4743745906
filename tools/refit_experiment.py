"""Dev-time experiment (CPU only, numpy): what would bottom-up refitted screen-space rectangles buy the ray caster?

    python tools/refit_experiment.py [n_tris] [width height]

Today project_kernel gives an inner child the rectangle of its projected 3-D BOX and a leaf child the (tighter)
rectangle of its projected TRIANGLE (rt_raycast.cu).  A refit would give an inner child the union of ITS children's
rectangles, recursively -- never larger, and ending in triangle-tight leaves.  This script rebuilds the same kind of
tree on the CPU (PLOC over Morton-sorted leaves, radius 32, as rt_bvh.cu), projects it for the lesson06 4K camera both
ways, and counts what an idealised traversal touches for every pixel of the traced rect: a child is entered when the
pixel lies in its rectangle and its nearest depth is not behind the ray's true hit (taken from the oracle), which is
what the real front-to-back traversal converges to.  Reported per variant: entered inner nodes and triangle tests per
traced ray, and the same per 8x4-pixel tile packet (the union over the tile's rays: what the warp actually executes).
`levels = k` refits only k levels up from the leaves (k cheap gather passes on the GPU instead of a dependent chain).
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def morton30(c):
    def spread(v):
        v = v.astype(np.uint32) & 0x3FF
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    q = np.clip((c * 1024.0).astype(np.int64), 0, 1023)
    return (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])


def half_area(lo, hi):
    d = hi - lo
    return d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]


def ploc(lo, hi, radius=32):
    """lo, hi: (T, 3) leaf boxes in Morton order.  Returns children (n_inner, 2) references (>= 0 inner id, < 0 ~leaf),
    child boxes clo/chi (n_inner, 2, 3), ids in creation order (children before parents), root id."""
    T = lo.shape[0]
    ref = ~np.arange(T, dtype=np.int64)
    lo, hi = lo.copy(), hi.copy()
    children = np.zeros((T - 1, 2), np.int64)
    clo, chi = np.zeros((T - 1, 2, 3)), np.zeros((T - 1, 2, 3))
    made = 0
    while ref.shape[0] > 1:
        n = ref.shape[0]
        best = np.full(n, np.inf)
        nn = np.full(n, -1, np.int64)
        for k in list(range(-radius, 0)) + list(range(1, radius + 1)):      # candidates in ascending j: ties go to the lower index
            if abs(k) >= n:
                continue
            i = np.arange(max(0, -k), min(n, n - k))
            j = i + k
            a = half_area(np.minimum(lo[i], lo[j]), np.maximum(hi[i], hi[j]))
            better = a < best[i]
            best[i[better]] = a[better]
            nn[i[better]] = j[better]
        idx = np.arange(n)
        mutual = (nn[nn] == idx) & (idx < nn)
        a_i, b_i = idx[mutual], nn[mutual]
        m = a_i.shape[0]
        ids = made + np.arange(m)
        children[ids, 0], children[ids, 1] = ref[a_i], ref[b_i]
        clo[ids, 0], chi[ids, 0], clo[ids, 1], chi[ids, 1] = lo[a_i], hi[a_i], lo[b_i], hi[b_i]
        made += m
        lo[a_i], hi[a_i] = np.minimum(lo[a_i], lo[b_i]), np.maximum(hi[a_i], hi[b_i])
        ref[a_i] = ids
        keep = np.ones(n, bool)
        keep[b_i] = False
        ref, lo, hi = ref[keep], lo[keep], hi[keep]
    assert made == T - 1
    return children, clo, chi, made - 1


def project_boxes(lo, hi, minv, o):
    """(n, 3) boxes -> rect (n, 4) = (sx_min, sx_max, sy_min, sy_max) and nearest depth (n,), from the 8 projected corners."""
    n = lo.shape[0]
    smin = np.full(n, np.inf); smax = np.full(n, -np.inf); tmin = np.full(n, np.inf); tmax = np.full(n, -np.inf); cmin = np.full(n, np.inf)
    for k in range(8):
        p = np.stack([hi[:, 0] if k & 1 else lo[:, 0], hi[:, 1] if k & 2 else lo[:, 1], hi[:, 2] if k & 4 else lo[:, 2]], 1) - o
        abc = p @ minv.T
        s, t = abc[:, 0] / abc[:, 2], abc[:, 1] / abc[:, 2]
        smin = np.minimum(smin, s); smax = np.maximum(smax, s); tmin = np.minimum(tmin, t); tmax = np.maximum(tmax, t)
        cmin = np.minimum(cmin, abc[:, 2])
    return np.stack([smin, smax, tmin, tmax], 1), cmin


def project_tris(P, minv, o):
    abc = (P - o) @ minv.T                                  # (T, 3 vertices, 3)
    s, t = abc[..., 0] / abc[..., 2], abc[..., 1] / abc[..., 2]
    return np.stack([s.min(1), s.max(1), t.min(1), t.max(1)], 1), abc[..., 2].min(1)


def main(n_tris=100_000, W=3840, H=2160):
    import rendering as ren
    import oracle
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes
    oracle.build()
    rows = scenes.dragon(n_tris)
    P = rows[:, :3].astype(np.float64).reshape(-1, 3, 3)
    T = P.shape[0]
    tlo, thi = P.min(1), P.max(1)
    slo, shi = tlo.min(0), thi.max(0)
    cen = ((tlo + thi) * 0.5 - slo) / (shi - slo).max()
    order = np.argsort(morton30(cen), kind="stable")
    t0 = time.time()
    children, clo, chi, root = ploc(tlo[order], thi[order])
    print(f"PLOC tree over {T} triangles in {time.time() - t0:.1f} s (numpy)")

    world, view, proj = scenes.lesson_camera(ren, 6, 0.5, W, H)
    cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
    o = cam[0:3].astype(np.float64)
    M = cam[3:12].astype(np.float64).reshape(3, 3).T        # columns U, V, W
    minv = np.linalg.inv(M)

    # per child: box rectangle; leaf children: intersected with the triangle's rectangle
    n_inner = T - 1
    rect = np.zeros((n_inner, 2, 4)); zmin = np.zeros((n_inner, 2))
    for c in range(2):
        rect[:, c], zmin[:, c] = project_boxes(clo[:, c], chi[:, c], minv, o)
    trect, tz = project_tris(P[order], minv, o)
    leaf = children < 0
    for c in range(2):
        sel = leaf[:, c]
        slot = ~children[sel, c]
        r = rect[sel, c]
        r[:, 0] = np.maximum(r[:, 0], trect[slot, 0]); r[:, 1] = np.minimum(r[:, 1], trect[slot, 1])
        r[:, 2] = np.maximum(r[:, 2], trect[slot, 2]); r[:, 3] = np.minimum(r[:, 3], trect[slot, 3])
        rect[sel, c] = r
        zmin[sel, c] = np.maximum(zmin[sel, c], tz[slot])

    def refit(levels):
        """levels = None: full bottom-up refit (ids are in creation order, children first); k: k gather passes."""
        r, z = rect.copy(), zmin.copy()
        if levels is None:
            for i in range(n_inner):
                for c in range(2):
                    ch = children[i, c]
                    if ch >= 0:
                        u = (min(r[ch, 0, 0], r[ch, 1, 0]), max(r[ch, 0, 1], r[ch, 1, 1]), min(r[ch, 0, 2], r[ch, 1, 2]), max(r[ch, 0, 3], r[ch, 1, 3]))
                        r[i, c] = (max(r[i, c, 0], u[0]), min(r[i, c, 1], u[1]), max(r[i, c, 2], u[2]), min(r[i, c, 3], u[3]))
                        z[i, c] = max(z[i, c], min(z[ch, 0], z[ch, 1]))
            return r, z
        for _ in range(levels):
            src_r, src_z = r.copy(), z.copy()
            for c in range(2):
                sel = children[:, c] >= 0
                ch = children[sel, c]
                r[sel, c, 0] = np.maximum(r[sel, c, 0], np.minimum(src_r[ch, 0, 0], src_r[ch, 1, 0]))
                r[sel, c, 1] = np.minimum(r[sel, c, 1], np.maximum(src_r[ch, 0, 1], src_r[ch, 1, 1]))
                r[sel, c, 2] = np.maximum(r[sel, c, 2], np.minimum(src_r[ch, 0, 2], src_r[ch, 1, 2]))
                r[sel, c, 3] = np.minimum(r[sel, c, 3], np.maximum(src_r[ch, 0, 3], src_r[ch, 1, 3]))
                z[sel, c] = np.maximum(z[sel, c], np.minimum(src_z[ch, 0], src_z[ch, 1]))
        return r, z

    # rays of the traced rect and their true hit depth
    srect, _ = project_boxes(slo[None], shi[None], minv, o)
    px0 = max(0, int(np.floor((srect[0, 0] + 1) * W / 2 - 0.5)) - 2) // 8 * 8
    px1 = min(W, (int(np.ceil((srect[0, 1] + 1) * W / 2 - 0.5)) + 2) // 8 * 8 + 8)
    py0 = max(0, int(np.floor((1 - srect[0, 3]) * H / 2 - 0.5)) - 2) // 4 * 4
    py1 = min(H, (int(np.ceil((1 - srect[0, 2]) * H / 2 - 0.5)) + 2) // 4 * 4 + 4)
    print(f"traced rect x {px0}..{px1} y {py0}..{py1}: {(px1 - px0) * (py1 - py0) / (W * H):.3f} of the frame")
    bvh = oracle.bvh_build(rows)

    def count(r, z, label):
        visits = tests = pk_visits = pk_tests = rays = 0
        band = 256
        for y0 in range(py0, py1, band):
            y1 = min(py1, y0 + band)
            xs, ys = np.meshgrid(np.arange(px0, px1), np.arange(y0, y1))
            xs, ys = xs.ravel(), ys.ravel()
            sx = ((xs + 0.5) * (2.0 / W) - 1.0).astype(np.float32).astype(np.float64)
            sy = (1.0 - (ys + 0.5) * (2.0 / H)).astype(np.float32).astype(np.float64)
            rr = oracle.primary_rays(cam, W, H, rect=(px0, y0, px1 - px0, y1 - y0))
            thit = oracle.bvh_raycast(bvh, rr)[0].astype(np.float64)
            tile = (ys // 4) * (W // 8) + xs // 8
            rays += xs.shape[0]
            fr_ray, fr_node = np.arange(xs.shape[0]), np.full(xs.shape[0], root)
            while fr_ray.shape[0]:
                visits += fr_ray.shape[0]
                pk_visits += np.unique(tile[fr_ray] * (2 * T) + fr_node).shape[0]
                nxt_ray, nxt_node = [], []
                for c in range(2):
                    rc, zc, ch = r[fr_node, c], z[fr_node, c], children[fr_node, c]
                    hit = (sx[fr_ray] >= rc[:, 0]) & (sx[fr_ray] <= rc[:, 1]) & (sy[fr_ray] >= rc[:, 2]) & (sy[fr_ray] <= rc[:, 3]) & (zc <= thit[fr_ray])
                    lf = hit & (ch < 0)
                    tests += int(lf.sum())
                    pk_tests += np.unique(tile[fr_ray[lf]] * (2 * T) + (~ch[lf])).shape[0]
                    inner = hit & (ch >= 0)
                    nxt_ray.append(fr_ray[inner]); nxt_node.append(ch[inner])
                fr_ray, fr_node = np.concatenate(nxt_ray), np.concatenate(nxt_node)
        tiles = rays / 32
        print(f"{label:28s} per ray: {visits / rays:6.2f} inner nodes, {tests / rays:5.2f} triangle tests | per tile packet: "
              f"{pk_visits / tiles:6.1f} node visits, {pk_tests / tiles:5.1f} leaf visits")
        return pk_visits / tiles, pk_tests / tiles

    base = count(rect, zmin, "box rectangles (today)")
    for lv in (1, 2, 4, None):
        r, z = refit(lv)
        got = count(r, z, f"refit, {'all' if lv is None else lv} level(s)")
        print(f"    -> packet node visits {got[0] / base[0] - 1:+.1%}, leaf visits {got[1] / base[1] - 1:+.1%}")
    oracle.bvh_free(bvh)


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    main(*(a[:1] or [100_000]), *(a[1:3] if len(a) >= 3 else (3840, 2160)))
