"""Synthetic stand-ins for the reference's missing models/dragon.obj (listed in .MISSING_LARGE_BLOBS).

`dragon(n_tris)`  closed, bumpy, genus-1 surface: a (2,3) torus-knot tube with seamless multi-octave
                  sinusoidal displacement (phases from numpy.random.default_rng(0)), emitted as a triangle
                  soup of MeshVertex rows with normalised finite-difference normals, then put through
                  load_obj's scalar min/max normalisation (rendering/_loaders.py:34-38) so positions sit in
                  [-0.5, 0.5] like a mesh loaded by the reference.  dragon(100_000) is SURVEY.md's dragon100k.
`instanced(...)`  SURVEY.md's dragon10M: a 10x10 grid of yawed, 0.1-scaled copies flattened to one soup (the
                  reference has no instancing).
`write_obj`       the same soup as a v/vn/f Wavefront file, to exercise load_obj.
`lesson_camera`   World/View/Proj of tutorials/lesson06 (:74-83) or lesson08/09 (:90-99) at time t.

All arrays are numpy float32 (n_vertices, 20) in MeshVertex order: P 0-2, N 4-6, C 8-9 (UV), T, B unused.
"""
import numpy as np


def _grid_dims(n_tris):
    """nu x nv quads, 2 triangles each, nu ~ 5*nv (long thin tube)."""
    nv = max(3, int(round(np.sqrt(n_tris / 10.0))))
    nu = max(3, n_tris // (2 * nv))
    return nu, nv


def _surface(nu, nv, seed=0):
    rng = np.random.default_rng(seed)
    u = (np.arange(nu) / nu)[:, None]           # along the knot
    v = (np.arange(nv) / nv)[None, :]           # around the tube
    phi = 2 * np.pi * u

    def curve(ph):
        r = np.cos(3 * ph) + 2.0
        return np.stack([r * np.cos(2 * ph), r * np.sin(2 * ph), -np.sin(3 * ph)], axis=-1)

    c = curve(phi)                              # (nu, 1, 3)
    h = 1e-4
    t = curve(phi + h) - curve(phi - h)
    t /= np.linalg.norm(t, axis=-1, keepdims=True)
    up = np.array([0.0, 0.0, 1.0])
    n1 = np.cross(t, up)
    n1 /= np.linalg.norm(n1, axis=-1, keepdims=True)
    n2 = np.cross(t, n1)
    # seamless fBm-like displacement: integer frequencies, random phases/orientations
    disp = np.zeros((nu, nv))
    amp = 0.16
    for octave in range(5):
        fu, fv = int(rng.integers(2, 6)) * 2 ** octave, int(rng.integers(1, 4)) * 2 ** min(octave, 3)
        disp += amp * np.sin(2 * np.pi * (fu * u + fv * v) + rng.uniform(0, 2 * np.pi))
        disp += 0.5 * amp * np.sin(2 * np.pi * (fu * u - fv * v) + rng.uniform(0, 2 * np.pi))
        amp *= 0.5
    radius = 0.42 * (1.0 + disp)
    ang = 2 * np.pi * v
    p = c + radius[..., None] * (np.cos(ang)[..., None] * n1 + np.sin(ang)[..., None] * n2)   # (nu, nv, 3)
    # normals from central differences on the periodic grid
    du = np.roll(p, -1, axis=0) - np.roll(p, 1, axis=0)
    dv = np.roll(p, -1, axis=1) - np.roll(p, 1, axis=1)
    n = np.cross(du, dv)
    n /= np.maximum(np.linalg.norm(n, axis=-1, keepdims=True), 1e-20)
    outward = p - c
    flip = np.sum(n * outward, axis=-1, keepdims=True) < 0
    n = np.where(flip, -n, n)
    uv = np.stack(np.broadcast_arrays(u, v), axis=-1)
    return p, n, uv


def normalise_like_load_obj(rows):
    """rendering/_loaders.py:34-38, float32, scalar min/max over all coordinates."""
    v_min = rows[:, 0:3].min()
    v_max = rows[:, 0:3].max()
    v_size = v_max - v_min
    rows[:, 0:3] = (rows[:, 0:3] - v_min) / v_size - v_size * 0.5 / v_size
    return rows


def dragon(n_tris=100_000, seed=0, normalise=True):
    """Triangle soup (3*n, 20) float32 with exactly nu*nv*2 <= n_tris triangles (== n_tris for 100_000)."""
    nu, nv = _grid_dims(n_tris)
    p, n, uv = _surface(nu, nv, seed)
    i = np.arange(nu)[:, None]
    j = np.arange(nv)[None, :]
    i1, j1 = (i + 1) % nu, (j + 1) % nv
    a, b = np.broadcast_arrays(i, j)
    corners = [(a, b), (np.broadcast_to(i1, a.shape), b), (np.broadcast_to(i1, a.shape), np.broadcast_to(j1, a.shape)),
               (a, b), (np.broadcast_to(i1, a.shape), np.broadcast_to(j1, a.shape)), (a, np.broadcast_to(j1, a.shape))]
    ii = np.stack([c[0] for c in corners], axis=-1).reshape(-1)
    jj = np.stack([c[1] for c in corners], axis=-1).reshape(-1)
    rows = np.zeros((ii.size, 20), dtype=np.float32)
    rows[:, 0:3] = p[ii, jj]
    rows[:, 4:7] = n[ii, jj]
    rows[:, 8:10] = uv[ii, jj]
    return normalise_like_load_obj(rows) if normalise else rows


def instanced(base_rows, grid=10, scale=0.1, seed=1):
    """grid x grid copies of a soup on the xz-plane, each scaled and yawed by rng(seed); flattened."""
    rng = np.random.default_rng(seed)
    out = np.zeros((base_rows.shape[0] * grid * grid, 20), dtype=np.float32)
    n = base_rows.shape[0]
    k = 0
    for gx in range(grid):
        for gz in range(grid):
            yaw = rng.uniform(0, 2 * np.pi)
            c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
            rot = np.array([[c, 0, -s], [0, 1, 0], [s, 0, c]], dtype=np.float32)
            blk = out[k * n:(k + 1) * n]
            blk[:, 0:3] = (base_rows[:, 0:3] @ rot) * np.float32(scale)
            blk[:, 0] += np.float32((gx + 0.5) / grid - 0.5)
            blk[:, 2] += np.float32((gz + 0.5) / grid - 0.5)
            blk[:, 4:7] = base_rows[:, 4:7] @ rot
            blk[:, 8:10] = base_rows[:, 8:10]
            k += 1
    return out


def instanced_device(ren, base_rows, grid=10, scale=0.1, seed=1):
    """instanced() built on the GPU, straight into a MeshVertex device buffer (torch as plumbing: no 2.4 GB host array, no
    2.4 GB upload for dragon10M).  Same instances, same rng draws; coordinates may differ from the numpy version in the last
    bit (the 3x3 rotation is written out instead of going through BLAS) -- bench.py uses this, the parity tests use instanced()."""
    import torch
    rng = np.random.default_rng(seed)
    n = base_rows.shape[0]
    vb = ren.create_buffer(n * grid * grid, ren.MeshVertex)
    out = vb.tensor().view(torch.float32).view(-1, 20)
    base = torch.from_numpy(np.ascontiguousarray(base_rows)).to(out.device)
    k = 0
    for gx in range(grid):
        for gz in range(grid):
            yaw = rng.uniform(0, 2 * np.pi)
            c, s = float(np.float32(np.cos(yaw))), float(np.float32(np.sin(yaw)))
            blk = out[k * n:(k + 1) * n]
            for o in (0, 4):            # positions, normals: row vector times [[c, 0, -s], [0, 1, 0], [s, 0, c]]
                x, y, z = base[:, o], base[:, o + 1], base[:, o + 2]
                f = float(np.float32(scale)) if o == 0 else 1.0
                blk[:, o] = (x * c + z * s) * f
                blk[:, o + 1] = y * f
                blk[:, o + 2] = (z * c - x * s) * f
            blk[:, 0] += float(np.float32((gx + 0.5) / grid - 0.5))
            blk[:, 2] += float(np.float32((gz + 0.5) / grid - 0.5))
            blk[:, 8:10] = base[:, 8:10]
            k += 1
    vb.device_written()
    return vb


def write_obj(path, rows, with_uv=False):
    """Write a soup as a Wavefront OBJ (one v/vn[/vt] per corner, f in file order)."""
    with open(path, "w") as fh:
        fh.write("# rendertoy_b200 synthetic mesh\no dragon\nusemtl dragon\n")
        for r in rows:
            fh.write(f"v {r[0]:.9g} {r[1]:.9g} {r[2]:.9g}\n")
        for r in rows:
            fh.write(f"vn {r[4]:.9g} {r[5]:.9g} {r[6]:.9g}\n")
        if with_uv:
            for r in rows:
                fh.write(f"vt {r[8]:.9g} {r[9]:.9g}\n")
        for t in range(rows.shape[0] // 3):
            a, b, c = 3 * t + 1, 3 * t + 2, 3 * t + 3
            if with_uv:
                fh.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
            else:
                fh.write(f"f {a}//{a} {b}//{b} {c}//{c}\n")


def lesson_camera(ren, lesson, t, width, height):
    """(World, View, Proj) exactly as the tutorial main loops build them (lesson06:74-83, lesson08:90-99)."""
    eye_z = 2 if lesson == 6 else 1.0
    world = ren.matmul(ren.scale(1.0), ren.rotate(t, ren.make_float3(0, 1, 0)))
    view = ren.look_at(ren.make_float3(0, 0.3, eye_z), ren.make_float3(0, 0, 0), ren.make_float3(0, 1, 0))
    proj = ren.perspective(aspect_ratio=width / height)
    return world, view, proj
