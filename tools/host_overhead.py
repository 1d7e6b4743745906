"""Dev-time: where does the host time of one raster frame go? (run on the GPU box)"""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendertoy_b200 import lessons, scenes
rows = scenes.dragon(100_000)
vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
with ren.mapped(vb) as m:
    m.view(np.float32).reshape(rows.shape)[:] = rows
pres = ren.create_presenter(1920, 1080)
raster, g = lessons.build_lesson08(ren, pres.get_render_target())
cams = [scenes.lesson_camera(ren, 8, 0.01 * k, 1920, 1080) for k in range(300)]
def frames(n):
    for k in range(n):
        lessons.set_transforms(ren, g, *cams[k])
        lessons.render_frame(ren, raster, vb)
frames(20); torch.cuda.synchronize()
t0 = time.perf_counter(); frames(300); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"tutorial calls: enqueue {1e6*(t1-t0)/300:.1f} us/frame, total {1e6*(t2-t0)/300:.1f} us/frame")
import ctypes
from rendering._raster import transforms48
g48 = [(ctypes.c_float * 48)(*transforms48(*c).tolist()) for c in cams]
def frames2(n):
    for k in range(n):
        raster.draw_frame(vb, None, g48[k])
frames2(20); torch.cuda.synchronize()
t0 = time.perf_counter(); frames2(300); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"draw_frame: enqueue {1e6*(t1-t0)/300:.1f} us/frame, total {1e6*(t2-t0)/300:.1f} us/frame")
pr = cProfile.Profile(); pr.enable(); frames(300); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
