"""Dev-time A/B: PLOC vs LBVH trees -- build time, frame time, traversal statistics (cfg4 frame)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster
from rendertoy_b200 import scenes
from tools.quick_raycast_bench import cam

def main(n_tris, w=3840, h=2160, frames=30, lesson=6):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    target = ren.create_image2d(w, h, ren._core.RGBA)
    cams = [cam(lesson, 0.1 * k, w, h) for k in range(frames)]
    for builder in ("lbvh", "ploc"):
        rc = Raycaster([ren.Mesh(vb, None)], builder=builder)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): rc._build_ads()
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) / 5 * 1e3
        for k in range(3): rc.render(target, cams[k])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in cams: rc.render(target, c)
        e1.record(); torch.cuda.synchronize()
        st = torch.zeros(3, dtype=torch.int64, device="cuda")
        rc.render(target, cams[0], stats=st); torch.cuda.synchronize()
        nn, kk, rr = [int(v) for v in st.cpu()]
        print(f"T={n_tris} lesson{lesson:02d} {builder}: build {build_ms:.2f} ms (host wall, incl. Python), frame {e0.elapsed_time(e1) / frames * 1e3:.1f} us, "
              f"node visits/ray {nn / rr:.1f}, tri tests/ray {kk / rr:.2f}", flush=True)

if __name__ == "__main__":
    main(100_000)
    main(100_000, lesson=8)
    main(1_000_000, frames=10)
