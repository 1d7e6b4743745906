"""Dev-time experiment (CPU only, numpy; companion of tools/refit_experiment.py, same idealised traversal): how are a tile
packet's node visits distributed over the screen size of the visited nodes?  Visits of nodes that cover many tiles are
work every one of those tiles repeats -- what a two-level traversal (a block walks the top of the tree once for its region
and hands its warps a frontier) would share.

    python tools/visit_histogram.py
"""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")
import numpy as np
import refit_experiment as R
import rendering as ren, oracle
from rendering._raycaster import camera_frame
from rendertoy_b200 import scenes
oracle.build()
n_tris, W, H = 100_000, 3840, 2160
rows = scenes.dragon(n_tris)
P = rows[:, :3].astype(np.float64).reshape(-1, 3, 3); T = P.shape[0]
tlo, thi = P.min(1), P.max(1); slo, shi = tlo.min(0), thi.max(0)
cen = ((tlo + thi) * 0.5 - slo) / (shi - slo).max()
order = np.argsort(R.morton30(cen), kind="stable")
children, clo, chi, root = R.ploc(tlo[order], thi[order])
world, view, proj = scenes.lesson_camera(ren, 6, 0.5, W, H)
cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
o = cam[0:3].astype(np.float64); M = cam[3:12].astype(np.float64).reshape(3, 3).T; minv = np.linalg.inv(M)
n_inner = T - 1
rect = np.zeros((n_inner, 2, 4)); zmin = np.zeros((n_inner, 2))
for c in range(2):
    rect[:, c], zmin[:, c] = R.project_boxes(clo[:, c], chi[:, c], minv, o)
trect, tz = R.project_tris(P[order], minv, o)
for c in range(2):
    sel = children[:, c] < 0; slot = ~children[sel, c]; r = rect[sel, c]
    r[:, 0] = np.maximum(r[:, 0], trect[slot, 0]); r[:, 1] = np.minimum(r[:, 1], trect[slot, 1])
    r[:, 2] = np.maximum(r[:, 2], trect[slot, 2]); r[:, 3] = np.minimum(r[:, 3], trect[slot, 3])
    rect[sel, c] = r; zmin[sel, c] = np.maximum(zmin[sel, c], tz[slot])
# full refit
r, z = rect.copy(), zmin.copy()
for i in range(n_inner):
    for c in range(2):
        ch = children[i, c]
        if ch >= 0:
            r[i, c] = (max(r[i, c, 0], min(r[ch, 0, 0], r[ch, 1, 0])), min(r[i, c, 1], max(r[ch, 0, 1], r[ch, 1, 1])),
                       max(r[i, c, 2], min(r[ch, 0, 2], r[ch, 1, 2])), min(r[i, c, 3], max(r[ch, 0, 3], r[ch, 1, 3])))
            z[i, c] = max(z[i, c], min(z[ch, 0], z[ch, 1]))
# node's own rect (union of its two child rects) in tile units
own = np.stack([np.minimum(r[:, 0, 0], r[:, 1, 0]), np.maximum(r[:, 0, 1], r[:, 1, 1]), np.minimum(r[:, 0, 2], r[:, 1, 2]), np.maximum(r[:, 0, 3], r[:, 1, 3])], 1)
tiles_w = (own[:, 1] - own[:, 0]) * W / 2 / 8; tiles_h = (own[:, 3] - own[:, 2]) * H / 2 / 4
area_tiles = np.maximum(tiles_w, 0) * np.maximum(tiles_h, 0)
px0, px1, py0, py1 = 1208, 2560, 380, 1924
bvh = oracle.bvh_build(rows)
bins = [0, 1, 4, 16, 64, 256, 1024, 1e9]
hist = np.zeros(len(bins) - 1); tiles_total = 0
for y0 in range(py0, py1, 256):
    y1 = min(py1, y0 + 256)
    xs, ys = np.meshgrid(np.arange(px0, px1), np.arange(y0, y1)); xs, ys = xs.ravel(), ys.ravel()
    sx = ((xs + 0.5) * (2.0 / W) - 1.0).astype(np.float32).astype(np.float64); sy = (1.0 - (ys + 0.5) * (2.0 / H)).astype(np.float32).astype(np.float64)
    thit = oracle.bvh_raycast(bvh, oracle.primary_rays(cam, W, H, rect=(px0, y0, px1 - px0, y1 - y0)))[0].astype(np.float64)
    tile = (ys // 4) * (W // 8) + xs // 8
    tiles_total += xs.shape[0] / 32
    fr_ray, fr_node = np.arange(xs.shape[0]), np.full(xs.shape[0], root)
    while fr_ray.shape[0]:
        u = np.unique(tile[fr_ray] * (2 * T) + fr_node) % (2 * T)
        hist += np.histogram(area_tiles[u], bins=bins)[0]
        nr, nn = [], []
        for c in range(2):
            rc, zc, ch = r[fr_node, c], z[fr_node, c], children[fr_node, c]
            hit = (sx[fr_ray] >= rc[:, 0]) & (sx[fr_ray] <= rc[:, 1]) & (sy[fr_ray] >= rc[:, 2]) & (sy[fr_ray] <= rc[:, 3]) & (zc <= thit[fr_ray])
            inner = hit & (ch >= 0); nr.append(fr_ray[inner]); nn.append(ch[inner])
        fr_ray, fr_node = np.concatenate(nr), np.concatenate(nn)
print("packet node visits by the visited node's rectangle area (in 8x4 tiles), full refit:")
for a, b, hcount in zip(bins[:-1], bins[1:], hist):
    print(f"  area {a:>6g} .. {b:<6g} tiles: {hcount / tiles_total:6.2f} visits per packet")
print("  total", hist.sum() / tiles_total)
