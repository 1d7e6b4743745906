"""Dev-time experiment: how much would a SAH-quality hierarchy buy over the LBVH?

    python tools/sah_experiment.py build 100000 variants/sah100k.npz     (CPU, anywhere)
    python tools/sah_experiment.py run 100000 variants/sah100k.npz       (GPU box)

`build` makes a full-sweep SAH binary BVH (one triangle per leaf) of scenes.dragon(n) in the library's node / leaf
layout (rt_bvh.cuh); `run` renders the 4K lesson06 frame with the library's own LBVH and again with the SAH arrays
patched into the Raycaster, printing time per frame and traversal statistics for both.
"""
import sys
import numpy as np
sys.path.insert(0, ".")
sys.setrecursionlimit(10000)


def build(n_tris, out):
    from rendertoy_b200 import scenes
    rows = scenes.dragon(n_tris)
    P = rows[:, :3].astype(np.float32).reshape(-1, 3, 3)
    T = P.shape[0]
    ext = np.float32((P.reshape(-1, 3).max(0) - P.reshape(-1, 3).min(0)).max())
    pad = np.float32(ext * np.float32(7.62939453125e-6))
    lo = P.min(1) - pad
    hi = P.max(1) + pad
    cen = (0.5 * (lo.astype(np.float64) + hi)).astype(np.float64)
    nodes = np.zeros((max(T - 1, 1), 16), dtype=np.float32)
    nodes_i = nodes.view(np.int32)
    order = []
    counter = [0]

    def area(l, h):
        d = np.maximum(h - l, 0)
        return d[..., 0] * d[..., 1] + d[..., 1] * d[..., 2] + d[..., 2] * d[..., 0]

    def split(ids):
        """returns (left ids, right ids)"""
        n = len(ids)
        if n == 2:
            return ids[:1], ids[1:]
        best = (np.inf, None, None)
        for ax in range(3):
            o = ids[np.argsort(cen[ids, ax], kind="stable")]
            l_lo = np.minimum.accumulate(lo[o], 0); l_hi = np.maximum.accumulate(hi[o], 0)
            r_lo = np.minimum.accumulate(lo[o][::-1], 0)[::-1]; r_hi = np.maximum.accumulate(hi[o][::-1], 0)[::-1]
            k = np.arange(1, n)
            cost = area(l_lo[:-1], l_hi[:-1]) * k + area(r_lo[1:], r_hi[1:]) * (n - k)
            j = int(np.argmin(cost))
            if cost[j] < best[0]:
                best = (cost[j], o, j + 1)
        _, o, j = best
        return o[:j], o[j:]

    def emit(ids):
        """returns child reference and box"""
        if len(ids) == 1:
            slot = len(order)
            order.append(int(ids[0]))
            return ~slot, lo[ids[0]], hi[ids[0]]
        me = counter[0]; counter[0] += 1
        a, b = split(ids)
        ra, alo, ahi = emit(a)
        rb, blo, bhi = emit(b)
        nodes[me, 0:4] = (alo[0], ahi[0], alo[1], ahi[1])
        nodes[me, 4:8] = (blo[0], bhi[0], blo[1], bhi[1])
        nodes[me, 8:12] = (alo[2], ahi[2], blo[2], bhi[2])
        nodes_i[me, 12] = ra; nodes_i[me, 13] = rb
        return me, np.minimum(alo, blo), np.maximum(ahi, bhi)

    emit(np.arange(T))
    order = np.array(order)
    tris = np.zeros((T, 12), dtype=np.float32)
    tris[:, 0:3] = P[order, 0]
    tris.view(np.uint32)[:, 3] = order.astype(np.uint32)
    tris[:, 4:7] = P[order, 1] - P[order, 0]
    tris[:, 8:11] = P[order, 2] - P[order, 0]
    np.savez(out, nodes=nodes, tris=tris)
    print("built", T, "triangles,", counter[0], "inner nodes ->", out)


def run(n_tris, path):
    import torch
    import rendering as ren
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import scenes
    from tools.quick_raycast_bench import cam
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    rc = Raycaster([ren.Mesh(vb, None)])
    w, h = 3840, 2160
    target = ren.create_image2d(w, h, ren._core.RGBA)
    hits = torch.empty((h * w, 4), dtype=torch.float32, device="cuda")
    cams = [cam(6, 0.1 * k, w, h) for k in range(20)]

    def measure(tag):
        for k in range(3): rc.render(target, cams[k])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in cams: rc.render(target, c)
        e1.record(); torch.cuda.synchronize()
        st = torch.zeros(3, dtype=torch.int64, device="cuda")
        rc.render(target, cams[0], stats=st, hits=hits, frame_size=(w, h))
        torch.cuda.synchronize()
        n, k, r = [int(v) for v in st.cpu()]
        print(f"{tag}: {e0.elapsed_time(e1) / len(cams) * 1e3:.1f} us/frame, node visits/ray {n / r:.1f}, tri tests/ray {k / r:.2f}")
        return hits.clone()

    ref = measure("LBVH")
    d = np.load(path)
    rc.nodes[:d["nodes"].nbytes].copy_(torch.from_numpy(d["nodes"].view(np.uint8).reshape(-1)))
    rc.tris[:d["tris"].nbytes].copy_(torch.from_numpy(d["tris"].view(np.uint8).reshape(-1)))
    got = measure("SAH ")
    print("hit records identical:", bool((ref.view(torch.int32) == got.view(torch.int32)).all()))


if __name__ == "__main__":
    (build if sys.argv[1] == "build" else run)(int(sys.argv[2]), sys.argv[3])
