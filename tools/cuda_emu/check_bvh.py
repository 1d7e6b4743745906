"""tools/cuda_emu/check_bvh.py -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    python tools/cuda_emu/check_bvh.py [n_tris]

Runs the source of csrc/rt_bvh.cu (bounds, Morton codes, radix sort, Karras hierarchy + refit, PLOC rounds) on CPU threads
(tools/cuda_emu), validates the resulting tree on the host (a proper binary tree over all leaves rooted at 0, every child box
containing what is below it, height within the traversal's stack bound), and then traces a frame through the emulated
rt_raycast.cu WITH that tree: hits must equal the oracle's brute force bit for bit.  Kernel logic only.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def validate_tree(nodes, tris, n, label):
    """as tests/test_raycast_gpu.py:_validate_tree, on host arrays; returns the height"""
    child = nodes.view(np.int32)[:, 12:14]
    assert ((child >= -n) & (child < max(n - 1, 1))).all(), f"{label}: child reference out of range"
    v0, v1, v2 = tris[:, 0:3], tris[:, 0:3] + tris[:, 4:7], tris[:, 0:3] + tris[:, 8:11]
    seen_leaf, seen_node, height = np.zeros(n, np.int32), np.zeros(max(n - 1, 1), np.int32), 0
    stack = [(0, 1, None)]
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
                continue
            if bound is not None:
                assert (lo >= bound[0]).all() and (hi <= bound[1]).all(), f"{label}: child box of node {node} sticks out of its parent's"
            if c < 0:
                s = ~c
                seen_leaf[s] += 1
                pts = np.stack([v0[s], v1[s], v2[s]])
                assert (pts.min(0) >= lo).all() and (pts.max(0) <= hi).all(), f"{label}: leaf {s} outside its box"
            else:
                stack.append((c, depth + 1, (lo, hi)))
    assert (seen_leaf == 1).all(), f"{label}: leaves not reached exactly once"
    assert n == 1 or (seen_node == 1).all(), f"{label}: unreachable inner nodes"
    return height


def main(n_tris=700, W=64, H=40, quick=False):
    import emu_build
    import oracle
    import rendering as ren
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes
    oracle.build()
    B = C.CDLL(emu_build.build("rt_bvh"))
    R = C.CDLL(emu_build.build("rt_raycast"))
    VP, I64, I32, FP = C.c_void_p, C.c_int64, C.c_int, C.POINTER(C.c_float)
    for f in ("rt_bvh_node_bytes", "rt_bvh_tri_bytes", "rt_bvh_scratch_bytes"):
        getattr(B, f).restype = I64
        getattr(B, f).argtypes = [I64]
    B.rt_bvh_build.argtypes = [VP, VP, I64, VP, VP, VP, I32, VP]
    B.rt_last_error.restype = C.c_char_p
    R.rt_raycast_primary.argtypes = [VP, VP, I64, VP, VP, VP, FP, I32, I32, I32, I32, I32, I32, I32, C.c_uint64, VP, VP, I64, VP,
                                     C.POINTER(C.c_int), I32, VP, VP]
    R.rt_raycast_view_node_bytes.restype = I64
    R.rt_raycast_view_node_bytes.argtypes = [I64]
    R.rt_last_error.restype = C.c_char_p
    ok = True
    world, view, proj = scenes.lesson_camera(ren, 6, 0.7, W, H)
    cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
    for label, rows, indexed in (("soup", scenes.dragon(n_tris), False), ("indexed, shuffled", scenes.dragon(max(n_tris // 2, 40), seed=4), True),
                                 ("duplicates (equal Morton codes)", np.concatenate([scenes.dragon(60)] * 4), False), ("one triangle", scenes.dragon(60)[:3], False)):
        T = rows.shape[0] // 3
        pos4 = np.ones((rows.shape[0], 4), np.float32); pos4[:, :3] = rows[:, 0:3]
        nrm4 = np.zeros((rows.shape[0], 4), np.float32); nrm4[:, :3] = rows[:, 4:7]
        idx = None
        if indexed:
            idx = np.arange(rows.shape[0], dtype=np.int32).reshape(-1, 3)[np.random.default_rng(2).permutation(T)].ravel().copy()
        rr = oracle.primary_rays(cam, W, H)
        ref = oracle.raycast_brute(rows, rr, indices=idx)
        for builder, bname in ((0, "LBVH (Karras)"), (1, "PLOC")):
            if quick and builder == 1 and label not in ("soup", "one triangle"):
                continue            # PLOC is ~500 launches per build: the CPU test suite keeps two of its four cases
            t0 = time.time()
            nodes = np.zeros(int(B.rt_bvh_node_bytes(T)) // 4, np.float32)
            tris = np.zeros(int(B.rt_bvh_tri_bytes(T)) // 4, np.float32)
            scratch = np.zeros(int(B.rt_bvh_scratch_bytes(T)), np.uint8)
            rc = B.rt_bvh_build(pos4.ctypes.data, None if idx is None else idx.ctypes.data, T, nodes.ctypes.data, tris.ctypes.data, scratch.ctypes.data, builder, None)
            assert rc == 0, B.rt_last_error()
            height = validate_tree(nodes.reshape(-1, 16)[:max(T - 1, 1)], tris.reshape(-1, 12)[:T], T, f"{label}/{bname}")
            assert height <= 64
            hits = np.full((W * H, 4), np.nan, np.float32)
            vnodes = np.zeros(int(R.rt_raycast_view_node_bytes(T)), np.uint8)
            rc = R.rt_raycast_primary(nodes.ctypes.data, tris.ctypes.data, T, pos4.ctypes.data, nrm4.ctypes.data, None if idx is None else idx.ctypes.data,
                                      cam.ctypes.data_as(FP), W, H, 0, 0, W, H, 8, 0, hits.ctypes.data, None, W, None, None, 0, vnodes.ctypes.data, None)
            assert rc == 0, R.rt_last_error()
            same = (np.array_equal(hits[:, 0].view(np.uint32), ref[0].view(np.uint32)) and np.array_equal(hits[:, 1].view(np.uint32), ref[1])
                    and np.array_equal(hits[:, 2].view(np.uint32), ref[2].view(np.uint32)) and np.array_equal(hits[:, 3].view(np.uint32), ref[3].view(np.uint32)))
            ok &= same
            print(f"{label:34s} T={T:5d} {bname:14s} tree valid, height {height:2d}; traced with it: hits {'==' if same else '!='} oracle ({time.time() - t0:.1f} s)", flush=True)
    print("BUILDERS OK" if ok else "BUILDER MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    sys.exit(main(*(a[:1] or [700])))
