"""Dev-time: host cost per frame of the tile-partition loop of one rank (render stripes + push_stripes), run on the GPU box."""
import cProfile, pstats, sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster, camera_frame
from rendertoy_b200 import scenes, parallel
W, H = 3840, 2160
rows = scenes.dragon(100_000)
vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
with ren.mapped(vb) as m:
    m.view(np.float32).reshape(rows.shape)[:] = rows
rc = Raycaster([ren.Mesh(vb, None)])
cams = []
for k in range(256):
    world, view, proj = scenes.lesson_camera(ren, 6, 2 * np.pi * k / 256, W, H)
    cams.append(camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4)))
store = parallel.FrameStore(16, W, H)
targets = [ren.create_image2d(W, H, ren._core.RGBA) for _ in range(8)]
streams = [torch.cuda.Stream() for _ in range(4)]
push = torch.cuda.Stream()
ev = [torch.cuda.Event() for _ in range(8)]
stripes = (64, 8, 3)
def frames(n):
    for f in range(n):
        st = streams[f % 4]
        torch.cuda.set_stream(st)
        content = rc.render(targets[f % 8], cams[f % 256], stripes=stripes)
        ev[f % 8].record(st)
        push.wait_event(ev[f % 8])
        store.push_stripes(f % 16, targets[f % 8].ptr, content, stripes, push.cuda_stream)
frames(64); torch.cuda.synchronize()
t0 = time.perf_counter(); frames(1024); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"tile loop of one rank of 8: enqueue {1e6*(t1-t0)/1024:.1f} us/frame, total {1e6*(t2-t0)/1024:.1f} us/frame")
pr = cProfile.Profile(); pr.enable(); frames(1024); pr.disable(); torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("cumulative").print_stats(25)
