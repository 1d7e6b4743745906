import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendertoy_b200 import lessons, scenes
rows = scenes.dragon(100_000)
vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
with ren.mapped(vb) as m:
    m.view(np.float32).reshape(rows.shape)[:] = rows
NS = 8
rs = []
for _ in range(NS):
    pres = ren.create_presenter(1920, 1080)
    rs.append(lessons.build_lesson08(ren, pres.get_render_target()))
cam = scenes.lesson_camera(ren, 8, 0.5, 1920, 1080)
for r, g in rs:
    lessons.set_transforms(ren, g, *cam); lessons.render_frame(ren, r, vb)
torch.cuda.synchronize()
streams = [torch.cuda.Stream() for _ in range(NS)]
def run(n, use_streams):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for i in range(n):
        r, g = rs[i % NS]
        if use_streams:
            with torch.cuda.stream(streams[i % NS]):
                lessons.render_frame(ren, r, vb)
        else:
            lessons.render_frame(ren, r, vb)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    return 1e6 * (t1 - t0) / n, 1e6 * (t2 - t0) / n
for us in (0, 1, 0, 1):
    print("streams" if us else "single ", "enqueue %.1f us/frame total %.1f us/frame" % run(400, us))
