"""Dev-time: time the raster frame (clear+clear+draw) for dragon(n) at WxH with CUDA events."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendertoy_b200 import lessons, scenes

def run(n_tris, w, h, lesson=8, frames=50):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    pres = ren.create_presenter(w, h)
    if lesson == 8:
        raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    else:
        tex = np.random.default_rng(3).integers(0, 256, size=(500, 500, 3), dtype=np.uint8)
        raster, g, _, _ = lessons.build_lesson09(ren, pres.get_render_target(), tex)
    for k in range(5):
        lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8, 0.1 * k, w, h))
        lessons.render_frame(ren, raster, vb)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cams = [scenes.lesson_camera(ren, 8, 0.1 * k, w, h) for k in range(frames)]
    t0 = time.perf_counter()
    e0.record()
    for k in range(frames):
        lessons.set_transforms(ren, g, *cams[k])
        lessons.render_frame(ren, raster, vb)
    e1.record(); torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / frames * 1e3
    ms = e0.elapsed_time(e1) / frames
    print(f"lesson{lesson:02d} T={rows.shape[0]//3} {w}x{h}: {ms*1e3:.1f} us/frame (device), {wall*1e3:.1f} us/frame (wall) -> {rows.shape[0]//3/ms/1e3:.1f} Mtris/s")

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "ncu":
        run(100_000, 1920, 1080, 8, frames=3)
        run(1_000_000, 1920, 1080, 8, frames=3)
        sys.exit(0)
    run(100_000, 1920, 1080, 8)
    run(100_000, 1920, 1080, 9)
    run(100_000, 640, 480, 8)
    run(1_000_000, 1920, 1080, 8)
