"""Dev-time: BASELINE configs[4] -- 10M-triangle instanced dragon, 1080p, raster + raycast, one GPU's share.
Checks raster-vs-oracle on one frame (depth bits), raster-vs-raycast visibility, and times both paths."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster, camera_frame
from rendertoy_b200 import lessons, scenes
import oracle

W, H = 1920, 1080
t0 = time.perf_counter()
base = scenes.dragon(100_000)
rows = scenes.instanced(base, grid=10, scale=0.1, seed=1)
print(f"scene: {rows.shape[0]//3} triangles generated in {time.perf_counter()-t0:.1f} s", flush=True)
vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
vb.set(rows.view(ren.MeshVertex).reshape(-1))
pres = ren.create_presenter(W, H)
raster, g = lessons.build_lesson08(ren, pres.get_render_target())
cam = scenes.lesson_camera(ren, 8, 0.5, W, H)
lessons.set_transforms(ren, g, *cam)
lessons.render_frame(ren, raster, vb)
torch.cuda.synchronize()
t0 = time.perf_counter()
res = oracle.draw_triangles(8, W, H, rows, lessons.globals_as_floats(g))
print(f"oracle frame: {time.perf_counter()-t0:.1f} s on {oracle.num_threads()} threads, stats {res.stats}", flush=True)
depth = raster.get_depth_buffer().get().reshape(H, W)
bgra = raster.get_render_target().get()
print("raster depth mismatches:", int((depth != res.depth).sum()), "colour mismatches outside ties:", int(((bgra != res.bgra).any(-1) & (res.tie == 0)).sum()))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
cams = [scenes.lesson_camera(ren, 8, 2 * np.pi * k / 256, W, H) for k in range(16)]
e0.record()
for k in range(16):
    lessons.set_transforms(ren, g, *cams[k]); lessons.render_frame(ren, raster, vb)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 16
print(f"raster: {ms*1e3:.0f} us/frame -> {rows.shape[0]//3/ms/1e3:.0f} Mtris/s")
torch.cuda.synchronize(); t0 = time.perf_counter()
rc = Raycaster([ren.Mesh(vb, None)])
torch.cuda.synchronize(); print(f"BVH build: {1e3*(time.perf_counter()-t0):.1f} ms")
world, view, proj = cam
cf = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
target = ren.create_image2d(W, H, ren._core.RGBA)
hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
rc.render(target, cf, hits=hits); torch.cuda.synchronize()
ids = hits.cpu().numpy()[:, 1].view(np.uint32).reshape(H, W)
ras = np.where(res.winner == 0xFFFFFFFF, 0xFFFFFFFF, res.winner // 2)
both = (ids != 0xFFFFFFFF) & (ras != 0xFFFFFFFF)
print(f"raycast vs raster: same triangle on {(ids == ras)[both].mean():.5f} of {int(both.sum())} commonly covered pixels; coverage agreement {((ids != 0xFFFFFFFF) == (ras != 0xFFFFFFFF)).mean():.5f}")
e0.record()
for k in range(16):
    world, view, proj = cams[k]
    rc.render(target, camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4)))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 16
print(f"raycast: {ms*1e3:.0f} us/frame -> {W*H/ms/1e3:.0f} Mrays/s")
