"""Dev-time: time BVH build and fused primary-ray render at 4K / 1080p."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster, camera_frame
from rendertoy_b200 import scenes

def cam(lesson, t, w, h):
    world, view, proj = scenes.lesson_camera(ren, lesson, t, w, h)
    return camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))

def run(n_tris, w, h, lesson, frames=20):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = Raycaster([ren.Mesh(vb, None)])
    torch.cuda.synchronize(); tb = time.perf_counter() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); rc._build_ads(); e1.record(); torch.cuda.synchronize()
    target = ren.create_image2d(w, h, ren._core.RGBA)
    cams = [cam(lesson, 0.1 * k, w, h) for k in range(frames)]
    for k in range(3): rc.render(target, cams[k])
    torch.cuda.synchronize()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for k in range(frames): rc.render(target, cams[k])
    f1.record(); torch.cuda.synchronize()
    ms = f0.elapsed_time(f1) / frames
    cov = (target.get()[:, :, 3] != 0).mean()
    print(f"T={rows.shape[0]//3} {w}x{h} lesson{lesson:02d}: build {e0.elapsed_time(e1)*1e3:.0f} us (first {tb*1e3:.1f} ms), render {ms*1e3:.1f} us/frame -> {w*h/ms/1e3:.1f} Mrays/s, coverage {cov:.3f}")

if __name__ == "__main__":
    import os
    if os.environ.get("RT_VIEW_REFIT"):      # experimental tightening passes (rt_raycast_set_view_refit), A/B by hand
        from rendertoy_b200 import _native
        _native.call("rt_raycast_set_view_refit", int(os.environ["RT_VIEW_REFIT"]))
        print("view refit passes:", os.environ["RT_VIEW_REFIT"])
    mode = sys.argv[1] if len(sys.argv) > 1 else ""
    fr = 3 if mode == "ncu" else (40 if mode == "ab" else 20)     # ncu: few launches to profile; ab: the two 4K frames only, 40 frames each
    run(100_000, 3840, 2160, 6, fr); sys.stdout.flush()
    run(100_000, 3840, 2160, 8, fr)
    if mode == "":
        run(100_000, 1920, 1080, 8, fr)
        run(1_000_000, 3840, 2160, 6, fr)
