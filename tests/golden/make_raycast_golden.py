"""Second, independent restatement of the ray-cast definition -- numpy float32, elementwise operations in the stated order
(numpy never fuses multiply-add) -- used to pin oracle/raycast_oracle.c and the CUDA path on a fixed scene.

    python tests/golden/make_raycast_golden.py        -> tests/golden/raycast_numpy_dragon300.npz

Definition (rendering/_raycaster.py:35-36 is `pass` in the reference, so this IS the specification):
ray (x, y) of a W x H frame: s = ((x + 0.5) * (2 / W)) - 1, t = 1 - ((y + 0.5) * (2 / H)), dir = (U s + V t) + W, origin
from the camera frame; Moller-Trumbore in float32 exactly as written below; closest hit = minimum over
(bits(t) << 32 | triangle id); Lambert shade d = sum_i max(0.2, N_i . l) w_i with l = (1, 1, 1) / sqrt(3), BGRA8 by
round-half-even of 255 d.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
f32 = np.float32


def primary_rays(cam, W, H):
    cam = np.asarray(cam, f32)
    xs = (np.arange(W, dtype=f32) + f32(0.5)) * (f32(2.0) / f32(W)) - f32(1.0)
    ys = f32(1.0) - (np.arange(H, dtype=f32) + f32(0.5)) * (f32(2.0) / f32(H))
    sx, sy = np.meshgrid(xs, ys)                                  # (H, W)
    d = [(cam[3 + a] * sx + cam[6 + a] * sy) + cam[9 + a] for a in range(3)]
    return cam[0:3], np.stack(d, -1).reshape(-1, 3).astype(f32)


def raycast(o, d, tri):
    """o (3,), d (R, 3), tri (T, 3, 3) float32 -> t, id, u, v per ray"""
    R = d.shape[0]
    best = np.full(R, np.uint64(0xFFFFFFFFFFFFFFFF))
    bu = np.zeros(R, f32); bv = np.zeros(R, f32)
    dx, dy, dz = d[:, 0], d[:, 1], d[:, 2]
    with np.errstate(all="ignore"):
        for k in range(tri.shape[0]):
            v0, v1, v2 = tri[k]
            e1 = v1 - v0; e2 = v2 - v0
            px = dy * e2[2] - dz * e2[1]; py = dz * e2[0] - dx * e2[2]; pz = dx * e2[1] - dy * e2[0]
            det = (e1[0] * px + e1[1] * py) + e1[2] * pz
            inv = f32(1.0) / det
            tx, ty, tz = o[0] - v0[0], o[1] - v0[1], o[2] - v0[2]
            u = ((tx * px + ty * py) + tz * pz) * inv
            qx = ty * e1[2] - tz * e1[1]; qy = tz * e1[0] - tx * e1[2]; qz = tx * e1[1] - ty * e1[0]
            v = ((dx * qx + dy * qy) + dz * qz) * inv
            t = ((e2[0] * qx + e2[1] * qy) + e2[2] * qz) * inv
            ok = (det != 0) & (u >= 0) & ~(u > 1) & (v >= 0) & ~(u + v > 1) & (t > 0) & (t != np.inf)
            key = (t.astype(f32).view(np.uint32).astype(np.uint64) << np.uint64(32)) | np.uint64(k)
            better = ok & (key < best)
            best = np.where(better, key, best); bu = np.where(better, u, bu); bv = np.where(better, v, bv)
    miss = best == np.uint64(0xFFFFFFFFFFFFFFFF)
    t = np.where(miss, np.uint32(0x7F800000), (best >> np.uint64(32)).astype(np.uint32)).astype(np.uint32).view(f32)
    ids = np.where(miss, np.uint32(0xFFFFFFFF), (best & np.uint64(0xFFFFFFFF)).astype(np.uint32)).astype(np.uint32)
    return t, ids, bu.astype(f32), bv.astype(f32)


def shade_lambert(nrm, ids, u, v):
    """nrm (T, 3, 3) float32 vertex normals -> (R, 4) BGRA8"""
    n = f32(0.57735026918962576)            # 1 / sqrt(3) rounded to float32, as RT_INV_SQRT3
    hit = ids != 0xFFFFFFFF
    N = nrm[np.where(hit, ids, 0)]
    tl = [(N[:, i, 0] * n + N[:, i, 1] * n) + N[:, i, 2] * n for i in range(3)]
    w0 = f32(1.0) - u - v
    d = (np.maximum(f32(0.2), tl[0]) * w0 + np.maximum(f32(0.2), tl[1]) * u) + np.maximum(f32(0.2), tl[2]) * v
    x = d * f32(255.0)
    b = np.where(x > 0, np.rint(np.minimum(x, f32(255.0))), 0).astype(np.uint8)
    out = np.stack([b, b, b, np.full_like(b, 255)], -1)
    out[~hit] = 0
    return out


def main():
    import rendering as ren
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes
    W, H = 48, 32
    rows = scenes.dragon(300)
    world, view, proj = scenes.lesson_camera(ren, 6, 0.5, W, H)
    cam = np.asarray(camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4)), f32)
    pos = rows[:, 0:3].astype(f32).reshape(-1, 3, 3)
    nrm = rows[:, 4:7].astype(f32).reshape(-1, 3, 3)
    o, d = primary_rays(cam, W, H)
    t, ids, u, v = raycast(o, d, pos)
    bgra = shade_lambert(nrm, ids, u, v).reshape(H, W, 4)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "raycast_numpy_dragon300.npz")
    np.savez_compressed(out, rows=rows, camera=cam, width=W, height=H, t=t, ids=ids, u=u, v=v, bgra=bgra)
    print(f"{out}: {pos.shape[0]} triangles, {W}x{H}, {(ids != 0xFFFFFFFF).sum()} hits")


if __name__ == "__main__":
    main()
