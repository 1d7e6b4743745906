"""Dev-time experiment (CPU only, numpy): the two-level packet traversal sketched in DESIGN.md section 8.

    python tools/region_experiment.py [n_tris] [region_tiles_x region_tiles_y] [a_max_tiles]

A block owns a region of RX x RY 8x4-pixel tiles.  Phase 1, once per region: walk the screen-space tree breadth first
with the REGION's rectangle (no depth culling: nothing is known yet) and stop at children that are leaves or whose
rectangle is at most `a_max` tiles in area -- the region's frontier.  Phase 2, per tile: the frontier entries overlapping
the tile are visited nearest first (entry test = rectangle + depth bound, per ray), and below each entered entry the usual
packet traversal runs.  Same idealised model as tools/refit_experiment.py (a child is entered when the pixel is inside its
rectangle and its depth bound is not behind the ray's true hit), converged refit rectangles.

Reported: phase-1 node tests per region, frontier size (mean / 99th percentile / max: sizes the shared-memory arrays),
candidates per tile, and per tile packet the node visits below the frontier + leaf visits, against the one-level walk.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def main(n_tris=100_000, RX=8, RY=4, a_max=8.0, W=3840, H=2160):
    import refit_experiment as R
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
    order = np.argsort(R.morton30(cen), kind="stable")
    children, clo, chi, root = R.ploc(tlo[order], thi[order])
    world, view, proj = scenes.lesson_camera(ren, 6, 0.5, W, H)
    cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
    o = cam[0:3].astype(np.float64)
    minv = np.linalg.inv(cam[3:12].astype(np.float64).reshape(3, 3).T)
    n_inner = T - 1
    r = np.zeros((n_inner, 2, 4)); z = np.zeros((n_inner, 2))
    for c in range(2):
        r[:, c], z[:, c] = R.project_boxes(clo[:, c], chi[:, c], minv, o)
    trect, tz = R.project_tris(P[order], minv, o)
    for c in range(2):
        sel = children[:, c] < 0
        slot = ~children[sel, c]
        q = r[sel, c]
        q[:, 0] = np.maximum(q[:, 0], trect[slot, 0]); q[:, 1] = np.minimum(q[:, 1], trect[slot, 1])
        q[:, 2] = np.maximum(q[:, 2], trect[slot, 2]); q[:, 3] = np.minimum(q[:, 3], trect[slot, 3])
        r[sel, c] = q
        z[sel, c] = np.maximum(z[sel, c], tz[slot])
    for i in range(n_inner):                                # converged refit (creation order: children first)
        for c in range(2):
            ch = children[i, c]
            if ch >= 0:
                r[i, c] = (max(r[i, c, 0], min(r[ch, 0, 0], r[ch, 1, 0])), min(r[i, c, 1], max(r[ch, 0, 1], r[ch, 1, 1])),
                           max(r[i, c, 2], min(r[ch, 0, 2], r[ch, 1, 2])), min(r[i, c, 3], max(r[ch, 0, 3], r[ch, 1, 3])))
                z[i, c] = max(z[i, c], min(z[ch, 0], z[ch, 1]))

    srect, _ = R.project_boxes(slo[None], shi[None], minv, o)
    px0 = max(0, int(np.floor((srect[0, 0] + 1) * W / 2 - 0.5)) - 2) // 8 * 8
    px1 = min(W, (int(np.ceil((srect[0, 1] + 1) * W / 2 - 0.5)) + 2) // 8 * 8 + 8)
    py0 = max(0, int(np.floor((1 - srect[0, 3]) * H / 2 - 0.5)) - 2) // 4 * 4
    py1 = min(H, (int(np.ceil((1 - srect[0, 2]) * H / 2 - 0.5)) + 2) // 4 * 4 + 4)
    bvh = oracle.bvh_build(rows)
    sxf = lambda x: (x + 0.5) * (2.0 / W) - 1.0
    syf = lambda y: 1.0 - (y + 0.5) * (2.0 / H)
    tile_area = (8 * 2.0 / W) * (4 * 2.0 / H)
    RW, RH = RX * 8, RY * 4

    phase1_tests, front_sizes, cand_per_tile = [], [], []
    visits = leafv = pk_visits = pk_leaf = tiles_n = 0
    one_pk = 0
    for ry0 in range(py0, py1, RH):
        ry1 = min(py1, ry0 + RH)
        xs, ys = np.meshgrid(np.arange(px0, px1), np.arange(ry0, ry1))
        xs, ys = xs.ravel(), ys.ravel()
        sx, sy = sxf(xs), syf(ys)
        thit = oracle.bvh_raycast(bvh, oracle.primary_rays(cam, W, H, rect=(px0, ry0, px1 - px0, ry1 - ry0)))[0].astype(np.float64)
        tile = (ys // 4) * (W // 8) + xs // 8
        tiles_n += xs.shape[0] / 32
        start_ray, start_ref = [], []
        for rx0 in range(px0, px1, RW):
            rx1 = min(px1, rx0 + RW)
            Rr = (sxf(rx0), sxf(rx1 - 1), syf(ry1 - 1), syf(ry0))            # region rectangle in NDC
            queue, front, tests = [root], [], 0
            while queue:                                                      # phase 1: breadth first, no depth culling
                nxt = []
                for nd in queue:
                    tests += 1
                    for c in range(2):
                        q = r[nd, c]
                        if q[0] <= Rr[1] and q[1] >= Rr[0] and q[2] <= Rr[3] and q[3] >= Rr[2]:
                            ch = children[nd, c]
                            if ch < 0 or (q[1] - q[0]) * (q[3] - q[2]) <= a_max * tile_area:
                                front.append((nd, c))
                            else:
                                nxt.append(ch)
                queue = nxt
            phase1_tests.append(tests)
            front_sizes.append(len(front))
            if not front:
                continue
            fn = np.array([f[0] for f in front]); fc = np.array([f[1] for f in front])
            fr, fz, fref = r[fn, fc], z[fn, fc], children[fn, fc]
            inreg = np.nonzero((xs >= rx0) & (xs < rx1))[0]
            # per tile: candidates = entries overlapping the tile's rectangle
            for ty in range(ry0, ry1, 4):
                for tx in range(rx0, rx1, 8):
                    tr = (sxf(tx), sxf(tx + 7), syf(ty + 3), syf(ty))
                    cand_per_tile.append(int(((fr[:, 0] <= tr[1]) & (fr[:, 1] >= tr[0]) & (fr[:, 2] <= tr[3]) & (fr[:, 3] >= tr[2])).sum()))
            # per ray: entered entries
            hit = ((sx[inreg, None] >= fr[None, :, 0]) & (sx[inreg, None] <= fr[None, :, 1]) & (sy[inreg, None] >= fr[None, :, 2])
                   & (sy[inreg, None] <= fr[None, :, 3]) & (fz[None, :] <= thit[inreg, None]))
            ri, ei = np.nonzero(hit)
            start_ray.append(inreg[ri]); start_ref.append(fref[ei])
        if not start_ray:
            continue
        fr_ray, fr_ref = np.concatenate(start_ray), np.concatenate(start_ref)
        lf = fr_ref < 0
        leafv += int(lf.sum())
        pk_leaf += np.unique(tile[fr_ray[lf]] * (2 * T) + (~fr_ref[lf])).shape[0]
        fr_ray, fr_node = fr_ray[~lf], fr_ref[~lf]
        while fr_ray.shape[0]:
            visits += fr_ray.shape[0]
            pk_visits += np.unique(tile[fr_ray] * (2 * T) + fr_node).shape[0]
            nr, nn = [], []
            for c in range(2):
                rc, zc, ch = r[fr_node, c], z[fr_node, c], children[fr_node, c]
                h = (sx[fr_ray] >= rc[:, 0]) & (sx[fr_ray] <= rc[:, 1]) & (sy[fr_ray] >= rc[:, 2]) & (sy[fr_ray] <= rc[:, 3]) & (zc <= thit[fr_ray])
                l2 = h & (ch < 0)
                leafv += int(l2.sum())
                pk_leaf += np.unique(tile[fr_ray[l2]] * (2 * T) + (~ch[l2])).shape[0]
                inner = h & (ch >= 0)
                nr.append(fr_ray[inner]); nn.append(ch[inner])
            fr_ray, fr_node = np.concatenate(nr), np.concatenate(nn)
    oracle.bvh_free(bvh)
    fs, p1, cp = np.array(front_sizes), np.array(phase1_tests), np.array(cand_per_tile)
    print(f"region {RX}x{RY} tiles, frontier at <= {a_max:g} tiles of area, {len(fs)} regions, {tiles_n:.0f} tiles")
    print(f"  phase 1: {p1.mean():.1f} node tests per region ({p1.mean() / (RX * RY):.2f} per tile), frontier size mean {fs.mean():.1f}, "
          f"99 % {np.percentile(fs, 99):.0f}, max {fs.max()}")
    print(f"  phase 2: {cp.mean():.2f} frontier candidates per tile (max {cp.max()}); per tile packet {pk_visits / tiles_n:.2f} node visits below the frontier, "
          f"{pk_leaf / tiles_n:.2f} leaf visits")


if __name__ == "__main__":
    a = sys.argv[1:]
    kw = {}
    if len(a) >= 3:
        kw.update(RX=int(a[1]), RY=int(a[2]))
    if len(a) >= 4:
        kw.update(a_max=float(a[3]))
    main(int(a[0]) if a else 100_000, **kw)
