"""Dev-time: SAH cost of the built tree (sum of inner-node areas / root area + leaf term) and cfg4 frame time, per builder."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster
from rendertoy_b200 import scenes
from tools.quick_raycast_bench import cam

def sah(rc):
    n = rc.n_triangles
    nd = rc.nodes.cpu().numpy().view(np.float32).reshape(-1, 16)[:n - 1]
    def area(lo, hi):
        d = np.maximum(hi - lo, 0); return d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    lo0 = nd[:, [0, 2, 8]]; hi0 = nd[:, [1, 3, 9]]; lo1 = nd[:, [4, 6, 10]]; hi1 = nd[:, [5, 7, 11]]
    a0, a1 = area(lo0, hi0), area(lo1, hi1)
    ch = nd.view(np.int32)[:, 12:14]
    root = area(np.minimum(lo0[:1], lo1[:1]), np.maximum(hi0[:1], hi1[:1]))[0]
    inner = (a0[ch[:, 0] >= 0].sum() + a1[ch[:, 1] >= 0].sum()) / root + 1.0
    leaf = (a0[ch[:, 0] < 0].sum() + a1[ch[:, 1] < 0].sum()) / root
    return inner, leaf

rows = scenes.dragon(100_000)
vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
with ren.mapped(vb) as m:
    m.view(np.float32).reshape(rows.shape)[:] = rows
w, h, frames = 3840, 2160, 20
target = ren.create_image2d(w, h, ren._core.RGBA)
cams = [cam(6, 0.1 * k, w, h) for k in range(frames)]
for builder in ("lbvh", "ploc"):
    rc = Raycaster([ren.Mesh(vb, None)], builder=builder)
    for k in range(3): rc.render(target, cams[k])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in cams: rc.render(target, c)
    e1.record(); torch.cuda.synchronize()
    i, l = sah(rc)
    print(f"{builder}: SAH inner {i:.1f} leaf {l:.1f}, frame {e0.elapsed_time(e1) / frames * 1e3:.1f} us", flush=True)
if len(sys.argv) > 1:
    d = np.load(sys.argv[1])
    rc.nodes[:d["nodes"].nbytes].copy_(torch.from_numpy(d["nodes"].view(np.uint8).reshape(-1)))
    rc.tris[:d["tris"].nbytes].copy_(torch.from_numpy(d["tris"].view(np.uint8).reshape(-1)))
    i, l = sah(rc)
    print(f"sweep SAH (CPU): SAH inner {i:.1f} leaf {l:.1f}")
