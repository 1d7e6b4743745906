"""Dev-time A/B: render() with per-frame screen-space nodes vs the 3-D slab traversal, same process, interleaved."""
import sys
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster
from rendertoy_b200 import scenes
from tools.quick_raycast_bench import cam

def main(n_tris=100_000, w=3840, h=2160, frames=40, lesson=6):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    rc = Raycaster([ren.Mesh(vb, None)])
    target = ren.create_image2d(w, h, ren._core.RGBA)
    cams = [cam(lesson, 0.1 * k, w, h) for k in range(frames)]
    for rep in range(2):
        for vn in (False, True):
            for k in range(3): rc.render(target, cams[k], view_nodes=vn)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for c in cams: rc.render(target, c, view_nodes=vn)
            e1.record(); torch.cuda.synchronize()
            print(f"T={n_tris} lesson{lesson:02d} view_nodes={vn}: {e0.elapsed_time(e1) / frames * 1e3:.1f} us/frame", flush=True)

if __name__ == "__main__":
    main()
    main(lesson=8)
    main(1_000_000, frames=20)
