"""Dev-time: launches every kernel of the library ONCE on the quoted configs inside cudaProfilerStart/Stop ranges, for one ncu run
(see tools/r02_profile.sh; an unbounded run over everything -- BVH builds launch ~250 kernels -- cost 40 GPU-minutes once):

    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:<not torch> -f -o /tmp/r02_all python tools/profile_all.py
    python tools/profile_inventory.py /tmp/r02_all.ncu-rep r02       # per-kernel summaries -> profiles/

`python tools/profile_all.py bvh` only builds the two trees (profiled with a -k filter and -c limit instead of ranges)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import rendering as ren
from rendering._raycaster import Raycaster, Ray, camera_frame
from rendertoy_b200 import lessons, scenes, _native

os.environ.pop("RENDERTOY_B200_GENERIC_RASTER", None)
BVH_ONLY = sys.argv[1:2] == ["bvh"]


class profiled:
    """with profiled(): ... -> the kernels launched inside are captured (ncu --profile-from-start off)"""

    def __enter__(self):
        torch.cuda.synchronize()
        torch.cuda.profiler.start()

    def __exit__(self, *a):
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
        return False



def cam12(lesson, t, w, h):
    world, view, proj = scenes.lesson_camera(ren, lesson, t, w, h)
    return camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))


def upload(rows):
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    return vb


rows = scenes.dragon(100_000)
vb = upload(rows)
W, H = 1920, 1080
if BVH_ONLY:
    Raycaster([ren.Mesh(vb, None)], builder="lbvh")      # first, so that a -c limit reaches karras / leaves_refit before PLOC's many rounds
    Raycaster([ren.Mesh(vb, None)], builder="ploc")
    torch.cuda.synchronize()
    sys.exit(0)
print("== raster lesson08 1080p 100k (cfg2): mesh_soa, fill_u64, raster_kernel<8,0>, coverage_kernel<8,0>, resolve_kernel<8>", flush=True)
r8, g8 = lessons.build_lesson08(ren, ren.create_presenter(W, H).get_render_target())
lessons.set_transforms(ren, g8, *scenes.lesson_camera(ren, 8, 0.3, W, H))
lessons.render_frame(ren, r8, vb)           # warm (the mesh_soa upload happens here, unprofiled; captured below on a second buffer)
lessons.set_transforms(ren, g8, *scenes.lesson_camera(ren, 8, 0.5, W, H))
with profiled():
    lessons.render_frame(ren, r8, vb)
vb2 = upload(rows[:30000])
with profiled():
    ren._core.mesh_soa(vb2)
print("== raster lesson09 1080p 100k (cfg3): raster_kernel<9,0>, resolve_kernel<9>", flush=True)
tex = np.random.default_rng(3).integers(0, 256, size=(500, 500, 3), dtype=np.uint8)
r9, g9, _, desc = lessons.build_lesson09(ren, ren.create_presenter(W, H).get_render_target(), tex)
lessons.set_transforms(ren, g9, *scenes.lesson_camera(ren, 8, 0.5, W, H))
with profiled():
    lessons.render_frame(ren, r9, vb)
print("== near-plane camera (large primitives through the work queue): coverage_kernel with work", flush=True)
from oracle import host_math as hm     # dev tool: host matrices only
near = tuple(ren.make_float4x4(np.ascontiguousarray(x)) for x in (hm.rotate(4.5, (0, 1, 0)), hm.look_at((0.12, 0.32, 0.3), (0, 0, 0), (0, 1, 0)), hm.perspective(aspect_ratio=W / H)))
lessons.set_transforms(ren, g8, *near)
with profiled():
    lessons.render_frame(ren, r8, vb)
print("== stripe partition (rank 3 of 8): fill_u64_owned, raster_kernel<8,1>, coverage_kernel<8,1>", flush=True)
r8.set_scissor(stripes=(64, 8, 3))
lessons.set_transforms(ren, g8, *scenes.lesson_camera(ren, 8, 0.5, W, H))
with profiled():
    lessons.render_frame(ren, r8, vb)
r8.set_scissor()
print("== clears and depth views: fill_u32, read_depth, write_depth", flush=True)
with profiled():
    ren.clear(r8.get_render_target()); r8.get_render_target().get()
    d = r8.get_depth_buffer().get(); r8.get_depth_buffer().set(d)
print("== draw_points: points_kernel<8>, resolve_points_kernel<8>", flush=True)
ren.clear(r8.get_render_target()); ren.clear(r8.get_depth_buffer(), 1.0)
with profiled():
    r8.draw_points(vb)
print("== user shaders through NVRTC: g_raster_triangles, g_resolve_triangles, g_raster_points, g_resolve_points", flush=True)
os.environ["RENDERTOY_B200_GENERIC_RASTER"] = "1"
rg, gg = lessons.build_lesson08(ren, ren.create_presenter(W, H).get_render_target())
os.environ.pop("RENDERTOY_B200_GENERIC_RASTER")
lessons.set_transforms(ren, gg, *scenes.lesson_camera(ren, 8, 0.5, W, H))
with profiled():
    lessons.render_frame(ren, rg, vb)
    ren.clear(rg.get_render_target()); ren.clear(rg.get_depth_buffer(), 1.0)
    rg.draw_points(vb)
print("== user kernel_main through NVRTC (lesson06 splat)", flush=True)


@ren.kernel_struct
class SplatT:
    World: ren.float4x4
    View: ren.float4x4
    Proj: ren.float4x4


@ren.kernel_main
def splat(im: ren.w_image2d_t, vertices: [ren.MeshVertex], info: SplatT):
    """
    int2 dim = get_image_dim(im);
    float3 P = vertices[thread_id].P;
    float3 C = vertices[thread_id].N * 0.5f + 0.5f;
    float4 H = (float4)(P.x, P.y, P.z, 1.0);
    H = mul(H, info.World);
    H = mul(H, info.View);
    H = mul(H, info.Proj);
    H.xyz /= H.w;
    if (any(H.xyz < (float3)(-1.0, -1.0, 0.0)) || any(H.xyz >= 1))
    return;
    int px = (int)(dim.x * (H.x * 0.5 + 0.5));
    int py = (int)(dim.y * (0.5 - H.y * 0.5));
    write_imagef(im, (int2)(px,py), (float4)(C.x, C.y, C.z, 1.0));
    """


ti = ren.create_struct(SplatT)
lessons.set_transforms(ren, ti, *scenes.lesson_camera(ren, 6, 0.5, 640, 480))
img = ren.create_image2d(640, 480, ren._core.RGBA)
with profiled():
    splat[vb.shape](img, vb, ti)

print("== BVH builds: bounds, morton, sort_hist/scan/scatter, ploc_* (PLOC) then karras, leaves_refit (LBVH)", flush=True)
rc = Raycaster([ren.Mesh(vb, None)], builder="ploc")
rcl = Raycaster([ren.Mesh(vb, None)], builder="lbvh")
torch.cuda.synchronize()
RW_, RH_ = 3840, 2160
print("== ray cast 4K lesson06 (cfg4): project_kernel, view_refit_kernel x2, raycast_kernel<8,0,0,1>", flush=True)
target = ren.create_image2d(RW_, RH_, ren._core.RGBA)
rc.render(target, cam12(6, 0.3, RW_, RH_))
with profiled():
    rc.render(target, cam12(6, 0.5, RW_, RH_))
print("== ray cast 4K lesson08 camera (frame-filling)", flush=True)
with profiled():
    rc.render(target, cam12(8, 0.5, RW_, RH_))
print("== stripe partition (rank 3 of 8), 4K lesson06", flush=True)
rc.render(target, cam12(6, 0.5, RW_, RH_), stripes=(64, 8, 3))       # (not captured: the same kernel on an eighth of the tiles)
torch.cuda.synchronize()
print("== ray cast 1080p textured (cfg3): raycast_kernel<9,...> with hits", flush=True)
t9 = ren.create_image2d(W, H, ren._core.RGBA)
hits = torch.empty((W * H, 4), dtype=torch.float32, device="cuda")
with profiled():
    rc.render(t9, cam12(8, 1.3, W, H), shader=_native.SHADER_LESSON09, texture_descriptor=desc, hits=hits)
print("== per-lane 3-D walk (view_nodes=False, FMA slab): raycast_kernel<8,0,1,0>", flush=True)
with profiled():
    rc.render(target, cam12(6, 0.5, RW_, RH_), view_nodes=False)
print("== generic rays: raycast_kernel<0,...>", flush=True)
n = 1 << 21
rng = np.random.default_rng(5)
rays = np.zeros((n, 8), np.float32)
rays[:, 0:3] = rng.uniform(-1.5, 1.5, (n, 3))
rays[:, 4:7] = rows[rng.integers(0, rows.shape[0], n), 0:3] - rays[:, 0:3]
rb = ren.create_buffer(n, Ray)
with ren.mapped(rb) as m:
    m.view(np.float32).reshape(n, 8)[:] = rays
with profiled():
    rc.ray_cast_native(rb)
print("done", flush=True)
