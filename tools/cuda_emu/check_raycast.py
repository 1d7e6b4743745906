"""tools/cuda_emu/check_raycast.py -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    python tools/cuda_emu/check_raycast.py [n_tris] [width height]

Runs the source of csrc/rt_raycast.cu on CPU threads (tools/cuda_emu) over a numpy-built PLOC tree in the library's node /
leaf layout and compares every hit record and shaded pixel with the oracle's brute-force definition, bit for bit, for: the
screen-space packet walk, the view-node refit passes, the per-lane 3-D walk, and the stripe partition of one frame over
several ranks (each rank one launch into the same frame; foreign rows must stay untouched).
This checks kernel LOGIC only -- it is how code written without GPU time left gets its first run.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, HERE)
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def library_bvh(rows):
    """PLOC tree (tools/refit_experiment.ploc) in rt_bvh.cuh layout: root 0, children with larger ids, leaf = ~slot."""
    import refit_experiment as R
    f32 = np.float32
    P = rows[:, :3].astype(f32).reshape(-1, 3, 3)
    T = P.shape[0]
    ext = f32((P.reshape(-1, 3).max(0) - P.reshape(-1, 3).min(0)).max())
    pad = f32(ext * f32(7.62939453125e-6))
    lo, hi = (P.min(1) - pad).astype(f32), (P.max(1) + pad).astype(f32)
    slo, shi = lo.min(0), hi.max(0)
    cen = ((lo.astype(np.float64) + hi) * 0.5 - slo) / float((shi - slo).max())
    order = np.argsort(R.morton30(cen), kind="stable")
    children, clo, chi, root = R.ploc(lo[order].astype(np.float64), hi[order].astype(np.float64))
    n_inner = T - 1
    assert root == n_inner - 1
    remap = lambda ref: np.where(ref >= 0, n_inner - 1 - ref, ref)
    nodes = np.zeros((n_inner, 16), f32)
    new = n_inner - 1 - np.arange(n_inner)
    nodes[new, 0], nodes[new, 1], nodes[new, 2], nodes[new, 3] = clo[:, 0, 0], chi[:, 0, 0], clo[:, 0, 1], chi[:, 0, 1]
    nodes[new, 4], nodes[new, 5], nodes[new, 6], nodes[new, 7] = clo[:, 1, 0], chi[:, 1, 0], clo[:, 1, 1], chi[:, 1, 1]
    nodes[new, 8], nodes[new, 9], nodes[new, 10], nodes[new, 11] = clo[:, 0, 2], chi[:, 0, 2], clo[:, 1, 2], chi[:, 1, 2]
    ni = nodes.view(np.int32)
    ni[new, 12], ni[new, 13] = remap(children[:, 0]), remap(children[:, 1])
    tris = np.zeros((T, 12), f32)
    tris[:, 0:3] = P[order, 0]
    tris.view(np.uint32)[:, 3] = order.astype(np.uint32)
    tris[:, 4:7] = P[order, 1] - P[order, 0]
    tris[:, 8:11] = P[order, 2] - P[order, 0]
    return nodes, tris


def rays_traced_tiles(cull, W, H):
    """number of 8x4 tiles the kernels trace (those touching the cull rect)"""
    if cull is None:
        return ((W + 7) // 8) * ((H + 3) // 4)
    x0, y0, x1, y1 = max(cull[0], 0), max(cull[1], 0), min(cull[2], W - 1), min(cull[3], H - 1)
    if x1 < x0 or y1 < y0:
        return 0
    return ((x1 >> 3) - (x0 >> 3) + 1) * ((y1 >> 2) - (y0 >> 2) + 1)


def main(n_tris=1500, W=96, H=64, quick=False):
    import emu_build
    import oracle
    import rendering as ren
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes
    oracle.build()
    L = C.CDLL(emu_build.build("rt_raycast"))
    VP, I64, I32, FP = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)
    L.rt_raycast_primary.argtypes = [VP, VP, I64, VP, VP, VP, FP, I32, I32, I32, I32, I32, I32, I32, C.c_uint64, VP, VP, I64, VP,
                                     C.POINTER(C.c_int), I32, VP, C.POINTER(C.c_int), VP]
    L.rt_raycast_view_node_bytes.restype = I64
    L.rt_raycast_view_node_bytes.argtypes = [I64]
    L.rt_last_error.restype = C.c_char_p
    counts = (C.c_ulonglong * 5)()

    rows = scenes.dragon(n_tris)
    T = rows.shape[0] // 3
    nodes, tris = library_bvh(rows)
    pos4 = np.ones((3 * T, 4), np.float32); pos4[:, :3] = rows[:, 0:3]
    nrm4 = np.zeros((3 * T, 4), np.float32); nrm4[:, :3] = rows[:, 4:7]
    vnodes = np.zeros(int(L.rt_raycast_view_node_bytes(T)), np.uint8)
    ok = True
    for lesson, t in (((8, 2.2),) if quick else ((6, 0.5), (8, 2.2))):
        world, view, proj = scenes.lesson_camera(ren, lesson, t, W, H)
        cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
        ref_t, ref_id, ref_u, ref_v = oracle.raycast_brute(rows, oracle.primary_rays(cam, W, H))
        ref_px = oracle.shade_hits(8, rows, ref_id, ref_u, ref_v).reshape(H, W, 4)
        cull = (C.c_int * 4)()
        lo, hi = rows[:, 0:3].min(0).astype(np.float64), rows[:, 0:3].max(0).astype(np.float64)
        L.rt_raycast_screen_bounds.argtypes = [FP, C.POINTER(C.c_double), C.POINTER(C.c_double), I32, I32, C.POINTER(C.c_int)]
        has_cull = L.rt_raycast_screen_bounds(cam.ctypes.data_as(FP), lo.ctypes.data_as(C.POINTER(C.c_double)), hi.ctypes.data_as(C.POINTER(C.c_double)), W, H, cull)
        variants = [("3-D per-lane walk", 0, 0.0, False), ("screen-space packets (measured path)", 0, 0.0, True)]
        variants += [(f"refit x{k}", k, 0.0, True) for k in (1, 4, 16)]
        # image-space partition: 3 ranks, stripes of 8 rows, each rank one launch into the same frame -> must equal the whole frame
        variants += [("3 ranks x stripes of 8 rows, packets", 0, 3, True), ("2 ranks x stripes of 16 rows, 3-D walk", 0, 2, False)]
        for label, passes, ranks, use_view in variants:
            assert L.rt_raycast_set_view_refit(passes) == 0
            hits = np.full((W * H, 4), np.nan, np.float32)
            bgra = np.full((H, W), 0x55555555, np.uint32)
            stats = np.zeros(3, np.uint64)
            t0 = time.time()
            L.emu_counts_read(counts, 1)
            for rank in range(max(ranks, 1)):
                before_h, before_c = hits.copy(), bgra.copy()
                stripes = (C.c_int * 3)(8 if ranks == 3 else 16, ranks, rank) if ranks else None
                rc = L.rt_raycast_primary(nodes.ctypes.data, tris.ctypes.data, T, pos4.ctypes.data, nrm4.ctypes.data, None, cam.ctypes.data_as(FP),
                                          W, H, 0, 0, W, H, 8, 0, hits.ctypes.data, bgra.ctypes.data, W, stats.ctypes.data,
                                          cull if has_cull else None, 0, vnodes.ctypes.data if use_view else None, stripes, None)
                assert rc == 0, L.rt_last_error()
                if ranks:   # a rank writes only rows of its own stripes
                    yy = np.arange(H)
                    foreign = np.repeat((yy // stripes[0]) % ranks != rank, W)
                    ok &= np.array_equal(hits.view(np.uint32)[foreign], before_h.view(np.uint32)[foreign]) and np.array_equal(bgra.ravel()[foreign], before_c.ravel()[foreign])
            L.emu_counts_read(counts, 1)
            tiles = max(rays_traced_tiles(cull if has_cull else None, W, H), 1)
            same = (np.array_equal(hits[:, 0].view(np.uint32), ref_t.view(np.uint32)) and np.array_equal(hits[:, 1].view(np.uint32), ref_id)
                    and np.array_equal(hits[:, 2].view(np.uint32), ref_u.view(np.uint32)) and np.array_equal(hits[:, 3].view(np.uint32), ref_v.view(np.uint32)))
            same_px = np.array_equal(bgra.view(np.uint8).reshape(H, W, 4), ref_px)
            ok &= same and same_px
            rays = max(int(stats[2]), 1)
            print(f"lesson{lesson:02d} {W}x{H} T={T}  {label:42s} hits {'==' if same else '!='} oracle, pixels {'==' if same_px else '!='}; "
                  f"{int(stats[0]) / rays:5.1f} node visits, {int(stats[1]) / rays:4.2f} triangle tests per ray; per traced tile: "
                  f"{counts[0] / tiles:6.1f} votes, {counts[1] / tiles:4.1f} warp-min, {counts[3] / 32 / tiles:6.1f} wide loads ({time.time() - t0:.1f} s)", flush=True)
    L.rt_raycast_set_view_refit(0)
    print("ALL BIT-EXACT" if ok else "MISMATCH")
    return 0 if ok else 1


def edges(quick=False):
    """Edge cases through every variant: the camera inside the mesh (unbounded rectangles), a single-triangle scene (the one-node tree with an empty second child), a sub-rectangle of the frame
    with a pitch, and no cull rectangle."""
    import emu_build
    import oracle
    from rendertoy_b200 import scenes
    oracle.build()
    L = C.CDLL(emu_build.build("rt_raycast"))
    VP, I64, I32, FP = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)
    L.rt_raycast_primary.argtypes = [VP, VP, I64, VP, VP, VP, FP, I32, I32, I32, I32, I32, I32, I32, C.c_uint64, VP, VP, I64, VP,
                                     C.POINTER(C.c_int), I32, VP, C.POINTER(C.c_int), VP]
    L.rt_raycast_view_node_bytes.restype = I64
    L.rt_raycast_view_node_bytes.argtypes = [I64]
    L.rt_last_error.restype = C.c_char_p
    W, H = (56, 36) if quick else (72, 44)
    ok = True
    rows_big = scenes.dragon(350 if quick else 900)
    one = rows_big[:3].copy()
    one[:, 0:3] = np.float32([[-0.3, -0.2, 0.1], [0.35, -0.25, 0.0], [0.05, 0.3, -0.1]])
    inside = np.array([0, 0, 0, 0.8, 0, 0, 0, 0.6, 0, 0, 0, 1], np.float32)
    outside = np.array([0.1, 0.2, -2.0, 0.5, 0.05, 0, 0, 0.4, 0.02, 0.03, -0.05, 1], np.float32)
    cases = [("camera inside the mesh", rows_big, inside, (0, 0, W, H)), ("single triangle", one, outside, (0, 0, W, H)),
             ("sub-rectangle of the frame", rows_big, outside, (8, 4, W - 19, H - 9)), ("dragon from outside, no cull rect", rows_big, outside, (0, 0, W, H))]
    for name, rows, cam, (x0, y0, w, h) in cases:
        T = rows.shape[0] // 3
        if T == 1:                                        # rt_bvh.cu: single_leaf_root_kernel
            f32 = np.float32
            P = rows[:, :3].astype(f32)
            pad = f32((P.max(0) - P.min(0)).max() * f32(7.62939453125e-6))
            lo, hi = P.min(0) - pad, P.max(0) + pad
            nodes = np.zeros((1, 16), f32)
            nodes[0, 0:4] = (lo[0], hi[0], lo[1], hi[1]); nodes[0, 4:8] = (np.inf, -np.inf, np.inf, -np.inf)
            nodes[0, 8:12] = (lo[2], hi[2], np.inf, -np.inf)
            nodes.view(np.int32)[0, 12:14] = (~0, ~0)
            tris = np.zeros((1, 12), f32)
            tris[0, 0:3] = P[0]; tris[0, 4:7] = P[1] - P[0]; tris[0, 8:11] = P[2] - P[0]
        else:
            nodes, tris = library_bvh(rows)
        pos4 = np.ones((3 * T, 4), np.float32); pos4[:, :3] = rows[:, 0:3]
        nrm4 = np.zeros((3 * T, 4), np.float32); nrm4[:, :3] = rows[:, 4:7]
        vnodes = np.zeros(int(L.rt_raycast_view_node_bytes(T)), np.uint8)
        rr = oracle.primary_rays(cam, W, H, rect=(x0, y0, w, h))
        ref_t, ref_id, ref_u, ref_v = oracle.raycast_brute(rows, rr)
        ref_px = oracle.shade_hits(8, rows, ref_id, ref_u, ref_v).reshape(h, w, 4)
        for label, passes, use_view in [("3-D", 0, False), ("packets", 0, True), ("refit x3", 3, True)]:
            L.rt_raycast_set_view_refit(passes)
            hits = np.full((w * h, 4), np.nan, np.float32)
            frame = np.full((H, W), 0x55555555, np.uint32)
            rc = L.rt_raycast_primary(nodes.ctypes.data, tris.ctypes.data, T, pos4.ctypes.data, nrm4.ctypes.data, None, cam.ctypes.data_as(FP),
                                      W, H, x0, y0, w, h, 8, 0, hits.ctypes.data, frame.ctypes.data + 4 * (y0 * W + x0), W, None, None, 0,
                                      vnodes.ctypes.data if use_view else None, None, None)
            assert rc == 0, L.rt_last_error()
            same = (np.array_equal(hits[:, 0].view(np.uint32), ref_t.view(np.uint32)) and np.array_equal(hits[:, 1].view(np.uint32), ref_id)
                    and np.array_equal(hits[:, 2].view(np.uint32), ref_u.view(np.uint32)) and np.array_equal(hits[:, 3].view(np.uint32), ref_v.view(np.uint32)))
            got_px = frame[y0:y0 + h, x0:x0 + w].copy().view(np.uint8).reshape(h, w, 4)
            untouched = frame.copy(); untouched[y0:y0 + h, x0:x0 + w] = 0x55555555
            same_px = np.array_equal(got_px, ref_px) and bool((untouched == 0x55555555).all())
            ok &= same and same_px
            print(f"{name:34s} {label:22s} hits {'==' if same else '!='} oracle, pixels {'==' if same_px else '!='} ({int((ref_id != 0xFFFFFFFF).sum())} of {w * h} rays hit)", flush=True)
    L.rt_raycast_set_view_refit(0)
    print("EDGE CASES BIT-EXACT" if ok else "EDGE CASE MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    if sys.argv[1:2] == ["edges"]:
        sys.exit(edges())
    a = [int(x) for x in sys.argv[1:]]
    sys.exit(main(*(a[:1] or [1500]), *(a[1:3] if len(a) >= 3 else (96, 64))))
