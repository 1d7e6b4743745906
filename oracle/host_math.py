"""oracle/host_math.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the reference's host camera/transform helpers (rendering/_core.py:421-548) with the
casting rules of the NumPy 1.x it was written for.  The reference itself raises under this image's
NumPy 2.3 (normalize -> make_float3 views 3 float64 as 6 float32), so these cannot be produced by
importing it; SURVEY.md Appendix D values are the regression anchors (parity unpinned for this
host math: no reference test, and the reference cannot run here).

Rules restated:
  * dot / cross  (_core.py:491-518): float32 products and sums, left to right, then .item().
  * normalize    (_core.py:503-510): l = sqrt(float64(dot)); float32 vector / float32(l)
                 (NumPy-1 value-based casting keeps `float32_array / float64_scalar` in float32).
  * rotate, perspective (_core.py:459-468, 540-548): float64 expressions rounded once to float32.
  * look_at      (_core.py:528-537): left-handed basis, rows = axes, last row = -dot(axis, eye).
All matrices are returned as (4, 4) float32, row-major, row-vector convention (v' = v @ M).
"""
import numpy as np

F = np.float32


def _v3(v):
    return np.asarray(v, dtype=np.float32).reshape(3)


def dot3(a, b):
    a, b = _v3(a), _v3(b)
    return float(F(F(F(a[0] * b[0]) + F(a[1] * b[1])) + F(a[2] * b[2])))


def cross3(a, b):
    a, b = _v3(a), _v3(b)
    return np.array([F(F(a[1] * b[2]) - F(a[2] * b[1])),
                     F(F(a[2] * b[0]) - F(a[0] * b[2])),
                     F(F(a[0] * b[1]) - F(a[1] * b[0]))], dtype=np.float32)


def normalize3(v):
    v = _v3(v)
    l = np.sqrt(np.float64(dot3(v, v)))
    return (v / F(l)).astype(np.float32)


def identity():
    return np.eye(4, dtype=np.float32)


def translate(x, y, z):
    m = np.eye(4, dtype=np.float64)
    m[3, :3] = (x, y, z)
    return m.astype(np.float32)


def scale(x, y=None, z=None):
    if y is None:
        y = z = x
    return np.diag(np.array([x, y, z, 1.0], dtype=np.float64)).astype(np.float32)


def rotate(angle, axis):
    c, s = np.cos(np.float64(angle)), np.sin(np.float64(angle))
    ax = _v3(axis)
    x, y, z = ax[0], ax[1], ax[2]          # float32 scalars; x*y etc. multiply in float32 first
    xx, yx, zx = np.float64(x * x), np.float64(y * x), np.float64(z * x)
    xy, yy, zy = np.float64(x * y), np.float64(y * y), np.float64(z * y)
    xz, yz, zz = np.float64(x * z), np.float64(y * z), np.float64(z * z)
    x, y, z = np.float64(x), np.float64(y), np.float64(z)
    m = np.array([
        [xx * (1 - c) + c, yx * (1 - c) + z * s, zx * (1 - c) - y * s, 0],
        [xy * (1 - c) - z * s, yy * (1 - c) + c, zy * (1 - c) + x * s, 0],
        [xz * (1 - c) + y * s, yz * (1 - c) - x * s, zz * (1 - c) + c, 0],
        [0, 0, 0, 1]], dtype=np.float64)
    return m.astype(np.float32)


def matmul(a, b):
    return (np.asarray(a, np.float32) @ np.asarray(b, np.float32)).astype(np.float32)


def look_at(camera, target, up):
    camera, target, up = _v3(camera), _v3(target), _v3(up)
    zaxis = normalize3((target - camera).astype(np.float32))
    xaxis = normalize3(cross3(up, zaxis))
    yaxis = cross3(zaxis, xaxis)
    m = np.zeros((4, 4), dtype=np.float32)
    m[0:3, 0], m[0:3, 1], m[0:3, 2] = xaxis, yaxis, zaxis
    m[3] = (-dot3(xaxis, camera), -dot3(yaxis, camera), -dot3(zaxis, camera), 1)
    return m


def perspective(fov=3.141593 / 4, aspect_ratio=1.0, znear=.01, zfar=100.0):
    hs = 1.0 / np.tan(np.float64(fov) / 2)
    ws = hs / aspect_ratio
    m = np.zeros((4, 4), dtype=np.float64)
    m[0, 0], m[1, 1] = ws, hs
    m[2, 2], m[2, 3] = zfar / (zfar - znear), 1.0
    m[3, 2] = -znear * zfar / (zfar - znear)
    return m.astype(np.float32)


def camera_frame(view, proj, world=None):
    """Primary-ray frame {origin, U, V, W} (12 float32) in MODEL space for the reference's camera
    convention: dir = (U*sx + V*sy) + W with sx,sy the NDC pixel centre.
    `world` must be a rigid rotation/translation (every tutorial uses rotate()); rays are moved to
    model space with its transpose/inverse in float64 and rounded once."""
    view = np.asarray(view, np.float64)
    proj = np.asarray(proj, np.float64)
    r = view[0:3, 0:3]                       # columns = xaxis, yaxis, zaxis
    eye = -(view[3, 0:3] @ np.linalg.inv(r))
    ws, hs = proj[0, 0], proj[1, 1]
    u, v, w = r[:, 0] / ws, r[:, 1] / hs, r[:, 2]
    if world is not None:
        wm = np.asarray(world, np.float64)
        winv = np.linalg.inv(wm)
        eye = (np.append(eye, 1.0) @ winv)[:3]
        u, v, w = (np.append(u, 0.0) @ winv)[:3], (np.append(v, 0.0) @ winv)[:3], (np.append(w, 0.0) @ winv)[:3]
    return np.concatenate([eye, u, v, w]).astype(np.float32)
