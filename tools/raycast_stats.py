"""Dev-time: traversal statistics (inner-node visits, triangle tests per traced ray) of the fused primary render."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster, camera_frame
from rendertoy_b200 import scenes
from tools.quick_raycast_bench import cam

def run(n_tris, w, h, lesson):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    rc = Raycaster([ren.Mesh(vb, None)])
    target = ren.create_image2d(w, h, ren._core.RGBA)
    for t in (0.0, 1.0, 2.5):
        st = torch.zeros(3, dtype=torch.int64, device="cuda")
        rc.render(target, cam(lesson, t, w, h), stats=st)
        torch.cuda.synchronize()
        n, k, r = [int(v) for v in st.cpu()]
        cov = float((target.get()[:, :, 3] != 0).mean())
        print(f"T={n_tris} {w}x{h} lesson{lesson:02d} t={t}: traced rays {r} ({r/(w*h):.3f} of frame), hit coverage {cov:.3f}, "
              f"node visits/ray {n/max(r,1):.1f}, tri tests/ray {k/max(r,1):.2f}")

if __name__ == "__main__":
    run(100_000, 3840, 2160, 6)
    run(1_000_000, 3840, 2160, 6)
