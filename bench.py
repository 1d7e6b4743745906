#!/usr/bin/env python
"""bench.py -- headline benchmark of rendertoy_b200 (contract: see DESIGN.md "Measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path raycast|raster]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

BASELINE.json's metric has two halves; each is measured on the config it is quoted on, one JSON line:

  primary    "Mrays/s closest-hit (dragon, 4K)"  -> configs[3]: dragon100k, 3840x2160, lesson06 camera orbit,
             fused primary rays + closest hit + Lambert shade, frames split across ranks, gathered to rank 0
  secondary  "Mtris/s raster"                    -> configs[1]: dragon100k, 1920x1080, lesson08 shaders,
             clear + clear + draw_triangles per frame (reported under "secondary" in the same line)

A step = FRAMES_PER_RANK frames per rank of the orbit animation (weak scaling: per-GPU work is fixed).
`value` = device-timed whole-job throughput with mesh/BVH resident; `e2e` = the same work driven through the
public `rendering` API with host-side inputs and every frame read back to pinned host memory.
`--impl reference` times the CPU oracle (oracle/: the restated reference pipeline; the reference has no ray
caster, so that half is our CPU BVH definition) on the host cores -- it is a baseline, not the target.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_TRIS = 100_000
RAY_W, RAY_H = 3840, 2160
RAS_W, RAS_H = 1920, 1080
FRAMES_PER_RANK = int(os.environ.get("RENDERTOY_B200_FRAMES_PER_RANK", "8"))
ORBIT = 256  # World = rotate(2*pi*k/256, y)  

METRIC_RAY = "Mrays/s closest-hit (dragon, 4K)"
METRIC_RAS = "Mtris/s raster"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """Per-launch DRAM bytes from the committed ncu --set full captures (profiles/traffic.json), or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    return json.load(open(p)) if os.path.exists(p) else {}


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while a timed region runs."""
    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
               0x100: "display_clock_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        while not self.stop_flag and self.nv is not None:
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------------
# scene helpers
# ---------------------------------------------------------------------------------------------------------

def orbit_t(k):
    return 2.0 * math.pi * (k % ORBIT) / ORBIT


def make_mesh(ren):
    from rendertoy_b200 import scenes
    rows = scenes.dragon(N_TRIS)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    return rows, vb


def ray_camera(ren, k):
    from rendering._raycaster import camera_frame
    from rendertoy_b200 import scenes
    world, view, proj = scenes.lesson_camera(ren, 6, orbit_t(k), RAY_W, RAY_H)
    return camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))


def dist_setup(n_gpus):
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def barrier_sync(world):
    import torch
    import torch.distributed as dist
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def all_ranks(ms, world):
    """every rank's value, in rank order (for the per-rank breakdown in the JSON line)"""
    import torch
    import torch.distributed as dist
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[dist.get_rank()] = ms
        dist.all_reduce(t)
        return [float(x) for x in t.cpu()]
    return [ms]


def max_over_ranks(ms, world):
    import torch
    import torch.distributed as dist
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return ms


# ---------------------------------------------------------------------------------------------------------
# primary: ray casting, configs[3]
# ---------------------------------------------------------------------------------------------------------

def bench_raycast(args, rank, world):
    import torch
    import rendering as ren
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import parallel
    from rendertoy_b200._native import SHADER_LESSON08

    rows, vb = make_mesh(ren)
    rc = Raycaster([ren.Mesh(vb, None)])
    F, n_frames = FRAMES_PER_RANK, FRAMES_PER_RANK * world
    my_frames = parallel.frame_indices(n_frames, rank, world)
    store = parallel.FrameStore(2 * n_frames, RAY_W, RAY_H) if (world > 1 and args.gather in ("peer", "copy")) else None
    fused = store is not None and store.ok and args.gather == "peer"
    pushed = store is not None and store.ok and args.gather == "copy"
    if pushed:  # ranks != 0 render locally and a copy engine pushes each finished frame into rank 0's frame store while the
        # next frame traces; rank 0 renders straight into the store
        slots2 = [[store.frame(b * n_frames + k) for k in my_frames] for b in range(2)]
        if rank == 0:
            targets2 = [[ren.Image(RAY_W, RAY_H, ren._core.RGBA, memory=m) for m in slots2[b]] for b in range(2)]
            targets = targets2[0]
        else:
            targets = [ren.create_image2d(RAY_W, RAY_H, ren._core.RGBA) for _ in range(F)]
        push_stream = torch.cuda.Stream()
        push_ptr = push_stream.cuda_stream
        rendered = [torch.cuda.Event() for _ in range(F)]
        local_ptrs = [t.ptr for t in targets] if rank != 0 else None
        local = gathered = None
    elif fused:   # render targets ARE rank 0's frame store (peer-mapped, double-buffered): the kernel's BGRA8 stores are the gather
        targets2 = [[ren.Image(RAY_W, RAY_H, ren._core.RGBA, memory=store.frame(b * n_frames + k)) for k in my_frames] for b in range(2)]
        targets = targets2[0]
        local = gathered = None
    else:
        targets = [ren.create_image2d(RAY_W, RAY_H, ren._core.RGBA) for _ in range(F)]     # F x 33 MB > L2
        local = torch.empty((F, RAY_H, RAY_W), dtype=torch.int32, device="cuda") if world > 1 else None
        gathered = torch.empty((n_frames, RAY_H, RAY_W), dtype=torch.int32, device="cuda") if (rank == 0 and world > 1) else None
    cams = {k: ray_camera(ren, k) for k in range(n_frames * (args.steps + args.warmup))}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(F * args.steps)]

    # frames of an orbit are independent: alternating them over a few streams overlaps one frame's tail (and the next
    # frame's projection pass) with its neighbour; each stream has its own screen-space node scratch inside Raycaster
    ray_streams = [torch.cuda.Stream() for _ in range(args.raycast_streams)] if args.raycast_streams > 1 else None
    main_stream = torch.cuda.current_stream()

    def step(s, timed_idx=None):
        tg = targets2[s % 2] if (fused or (pushed and rank == 0)) else targets
        if ray_streams is not None:
            for st in ray_streams:
                st.wait_stream(main_stream)
        for j, k in enumerate(my_frames):
            if ray_streams is not None:
                torch.cuda.set_stream(ray_streams[j % len(ray_streams)])
            if timed_idx is not None:
                ev[timed_idx * F + j][0].record()
            content = rc.render(tg[j], cams[s * n_frames + k])
            if timed_idx is not None:
                ev[timed_idx * F + j][1].record()
            if pushed and rank != 0:
                # the frame is the clear colour outside `content` (the scene's projected bounds): only that rect travels
                rendered[j].record()
                push_stream.wait_event(rendered[j])
                pushed_bytes[0] += store.push((s % 2) * n_frames + k, local_ptrs[j], content if args.sparse else full_rect, push_ptr)
        if ray_streams is not None:
            torch.cuda.set_stream(main_stream)
            for st in ray_streams:
                main_stream.wait_stream(st)
        if pushed:
            main_stream.wait_stream(push_stream)
        collect()

    full_rect = (0, 0, RAY_W - 1, RAY_H - 1)
    pushed_bytes = [0]

    def collect():   # the only collective: finished frames -> rank 0
        if fused or pushed:
            store.commit()
        elif world > 1:
            for j in range(F):
                local[j].copy_(targets[j].buffer.tensor().view(torch.int32).view(RAY_H, RAY_W))
            parallel.gather_frames(local, gathered, n_frames)

    for s in range(args.warmup):
        step(s)
    sampler = ClockSampler(torch.cuda.current_device()); sampler.start()
    barrier_sync(world)
    pushed_bytes[0] = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step(args.warmup + s, timed_idx=s)
    e1.record()
    barrier_sync(world)
    clocks = sampler.result()
    # untimed check of the gather: every rank compares the frames of the last step it rendered locally with what now
    # sits in its slots of rank 0's frame store (read back over NVLink), bit for bit
    gather_ok = None
    if pushed:
        s_last = args.warmup + args.steps - 1
        ok = 1
        if rank != 0:
            for j, k in enumerate(my_frames):
                slot = store.frame((s_last % 2) * n_frames + k).view(torch.int32)
                ok &= int(torch.equal(slot, targets[j].buffer.tensor().view(torch.int32).view(-1)))
        flag = torch.tensor([ok], dtype=torch.int32, device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        gather_ok = bool(flag.item())
        assert gather_ok, "frames in rank 0's frame store differ from the frames the ranks rendered"
    push_bytes_step = all_ranks(pushed_bytes[0] / max(args.steps, 1), world)
    ms_ranks = all_ranks(e0.elapsed_time(e1), world)
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    rays_total = RAY_W * RAY_H * n_frames * args.steps
    value = rays_total / (ms * 1e-3) / 1e6
    launch_ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))   # per frame, on its own stream, inside the timed region
    # with frames overlapping on several streams a launch's own duration includes its neighbours' share of the GPU:
    # the duration that explains `value` is the timed region divided by the launches in it
    kernel_ms = (e0.elapsed_time(e1) / (args.steps * F)) if ray_streams is not None else launch_ms
    # the same launch pair alone on the GPU (after the timed region)
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(F)]
    for j, k in enumerate(my_frames):
        iso[j][0].record(); rc.render(targets[j], cams[k]); iso[j][1].record()
    torch.cuda.synchronize()
    isolated_ms = float(np.mean([a.elapsed_time(b) for a, b in iso]))

    # instrumented pass (outside the timed region): node visits / triangle tests per ray
    stats = torch.zeros(3, dtype=torch.int64, device="cuda")
    rc.render(targets[0], cams[my_frames[0]], stats=stats)
    torch.cuda.synchronize()
    nodes, tests, rays = (int(x) for x in stats.cpu())

    # ---- e2e: public API, host inputs, every frame read back to pinned host memory
    host = [torch.zeros((RAY_H, RAY_W), dtype=torch.int32).pin_memory() for _ in range(2)]   # cleared, like the frames
    copy_stream = torch.cuda.Stream()
    copy_ptr = copy_stream.cuda_stream
    done = [torch.cuda.Event() for _ in range(2)]
    reader = parallel.SparseFrameCopier(RAY_W, RAY_H)
    from rendertoy_b200 import scenes
    from rendering._raycaster import camera_frame

    # e2e delivers frames to HOST memory: on one node every rank reads its own frames back over its own PCIe link, so
    # the GPU-side gather to rank 0 is not on this path (local targets, no collective)
    e2e_targets = targets if not fused else [ren.create_image2d(RAY_W, RAY_H, ren._core.RGBA) for _ in range(F)]

    # like the device-resident loop, the frames of a batch alternate over streams, but two of them: measured 43.6 Grays/s
    # with 2, 40.3 with 1, 38.8 with 4, 35.0 with 8 (the renders compete with the copy engine's reads)
    e2e_streams = ray_streams[:2] if ray_streams is not None else None

    def e2e_step(s):
        ray_streams = e2e_streams
        if ray_streams is not None:
            for st in ray_streams:
                st.wait_stream(main_stream)
        for j, k in enumerate(my_frames):
            world_m, view, proj = scenes.lesson_camera(ren, 6, orbit_t(s * n_frames + k), RAY_W, RAY_H)   # host inputs
            cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world_m, dtype=ren.float4x4))
            if ray_streams is not None:
                torch.cuda.set_stream(ray_streams[j % len(ray_streams)])
            content = rc.render(e2e_targets[j], cam)
            done[j % 2].record()
            copy_stream.wait_event(done[j % 2])
            # the frame is the clear colour outside `content`, and so is the (initially cleared) host frame outside the
            # content of the frame it held before: one pitched D2H copy of the union makes the host frame complete
            reader.copy(j % 2, host[j % 2].data_ptr(), e2e_targets[j].ptr, content if args.sparse_readback else full_rect, copy_ptr)
        if ray_streams is not None:
            torch.cuda.set_stream(main_stream)
            for st in ray_streams:
                main_stream.wait_stream(st)
        main_stream.wait_stream(copy_stream)

    e2e_step(0)
    barrier_sync(world)
    k_e2e = max(2, min(args.steps, 5))
    reader.bytes_moved = 0
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for s in range(k_e2e):
        e2e_step(1 + s)
    g1.record()
    barrier_sync(world)
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), world)
    # untimed check: the host frame that received the last frame is that frame, every pixel
    j_last = F - 1
    e2e_ok = bool(torch.equal(host[j_last % 2], e2e_targets[j_last].buffer.tensor().view(torch.int32).view(RAY_H, RAY_W).cpu()))
    assert e2e_ok, "sparse read-back: the host frame differs from the device frame"
    e2e_value = RAY_W * RAY_H * n_frames * k_e2e / (e2e_ms * 1e-3) / 1e6

    hbm_peak, peak_src = peaks()
    alg_bytes = 12 * RAY_W * RAY_H + N_TRIS * 96 + (2 * N_TRIS - 1) * 32
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    sm_count = torch.cuda.get_device_properties(0).multi_processor_count
    fp32_peak = sm_count * 128 * 2 * (clocks["sm_max_mhz"] or 1965) * 1e6 / 1e12
    flops = nodes * 10 + tests * 45              # totals of one full instrumented frame (culled pixels trace nothing)
    out = {
        "metric": METRIC_RAY, "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[3]: dragon100k (synthetic stand-in for the missing dragon.obj, 100000 triangles) "
                               "raycast 3840x2160, lesson06 camera orbit, primary rays + closest hit + Lambert shade",
                   "frames_per_rank_per_step": F,
                   "partition": "frames k = rank (mod N); " + ("every rank's kernel stores its pixels straight into rank 0's frame store "
                                "over NVLink (CUDA IPC peer memory, double-buffered), one stream-ordered 4-byte all-reduce per step" if fused else
                                "ranks != 0 render locally and push each finished frame into rank 0's frame store (CUDA IPC peer memory, "
                                "double-buffered, cleared at start) with an async pitched copy-engine transfer that overlaps the next frame"
                                + ("; only the pixel rect that can differ from the slot's content travels (union of the scene's projected "
                                   "bounds of this frame and of the slot's previous frame: the frame is the clear colour elsewhere)" if args.sparse else "")
                                + "; rank 0 renders in place; one stream-ordered 4-byte all-reduce per step" if pushed else
                                "framebuffers gathered to rank 0 with NCCL send/recv" if world > 1 else "single GPU, no gather"),
                   "l2": "each rank cycles 8 distinct 33 MB frame targets (265 MB > L2); mesh + BVH (~21 MB) stay "
                         "L2-resident by design, as they are reused every frame",
                   "streams": f"frames alternate over {args.raycast_streams} CUDA streams" if ray_streams else "single stream",
                   "bvh_build_excluded": True, "timed_region_ms_per_rank": ms_ranks,
                   **({"gather_verified": "every rank's locally rendered frames of the last step == its slots of rank 0's frame store, bit for bit",
                       "gather_bytes_per_step_per_rank": push_bytes_step, "full_frame_bytes_per_step_per_rank": 4 * RAY_W * RAY_H * F}
                      if gather_ok else {})},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": ncu_traffic().get("raycast_kernel"), "peak_source": peak_src,
                     "kernel": "raycast_kernel<8> (+ project_kernel, same launch pair)", "kernel_ms": kernel_ms, "kernel_ms_alone": isolated_ms, "kernel_ms_overlapped_launch": launch_ms,
                     "kernel_ms_note": "kernel_ms = timed region / launches in it (frames overlap on the streams named in config); "
                                       "kernel_ms_alone = one frame's launch pair with nothing else on the GPU; "
                                       "kernel_ms_overlapped_launch = CUDA events around each launch pair inside the timed region",
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "note": "HBM does not bind this kernel (BVH and its per-frame screen-space copy are L2-resident); the binding "
                             "unit is the instruction issue rate (compares, votes, branches of the packet walk: ~73 % of issue "
                             "slots busy in profiles/), see fp32 for the arithmetic it amounts to",
                     "fp32": {"achieved_tflops": flops / (kernel_ms * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                              "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp32_peak,
                              "inner_node_visits_per_ray": nodes / (RAY_W * RAY_H), "triangle_tests_per_ray": tests / (RAY_W * RAY_H),
                              "rays_traced_fraction": rays / (RAY_W * RAY_H),
                              "flop_model": "per voting lane: 10 float compares per inner node (two screen rectangles + depth bound "
                                            "each) + 45 flop per Moller-Trumbore test; peak counts FMA as 2 flop, this kernel is "
                                            "compiled -fmad=false for bit-exact parity"}},
        "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 192 * F, "d2h_bytes_per_step": reader.bytes_moved // k_e2e,
                "frame_bytes_per_step": 4 * RAY_W * RAY_H * F, "readback_verified": e2e_ok,
                "steps": k_e2e, "note": "per frame: host matrices -> camera frame -> rt_raycast_primary -> async pitched D2H copy into a "
                                        "pinned, initially cleared host frame" + (" of the pixel rect that can differ from the clear colour (union "
                                        "of the scene's projected bounds of this frame and of the frame the host buffer held before); the host "
                                        "frame is complete and checked against the device frame after the timed region" if args.sparse_readback
                                        else " of the whole 33 MB frame") + "; every rank reads back its own frames"},
        "gpu_launches": (2 + args.view_refit) * F * args.steps, "clocks": clocks,   # project_kernel (+ experimental refit passes) + raycast kernel per frame
    }
    return out, rows


# ---------------------------------------------------------------------------------------------------------
# secondary: rasterization, configs[1]
# ---------------------------------------------------------------------------------------------------------

def bench_raster(args, rank, world, rows=None):
    import torch
    import rendering as ren
    from rendertoy_b200 import lessons, scenes, parallel

    if rows is None:
        rows, vb = make_mesh(ren)
    else:
        vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
        with ren.mapped(vb) as m:
            m.view(np.float32).reshape(rows.shape)[:] = rows
    F, n_frames = FRAMES_PER_RANK, FRAMES_PER_RANK * world
    my_frames = parallel.frame_indices(n_frames, rank, world)
    # F independent targets (key 16.6 MB + colour 8.3 MB + records 12.8 MB each: ~300 MB > L2)
    store = parallel.FrameStore(2 * n_frames, RAS_W, RAS_H) if (world > 1 and args.gather != "nccl") else None
    # --raster-gather copy (EXPERIMENTAL, not yet measured): as for the ray-cast frames, ranks != 0 render locally and a copy engine
    # pushes Raster.content_rect into rank 0's frame store; rank 0 renders in place.  Default: every rank's kernels store into the store.
    ras_push = store is not None and store.ok and args.raster_gather == "copy"
    fused = store is not None and store.ok and not ras_push
    in_store = fused or (ras_push and rank == 0)         # this rank's render targets are slots of the store
    rasters, rasters_b = [], []
    for j in range(F):
        target = ren.Image(RAS_W, RAS_H, ren._core.RGBA, memory=store.frame(my_frames[j])) if in_store else \
            ren.create_presenter(RAS_W, RAS_H).get_render_target()
        rasters.append(lessons.build_lesson08(ren, target))
        if in_store:   # second half of the double-buffered frame store
            rasters_b.append(lessons.build_lesson08(ren, ren.Image(RAS_W, RAS_H, ren._core.RGBA, memory=store.frame(n_frames + my_frames[j]))))
    if ras_push:
        push_stream = torch.cuda.Stream()
        push_ptr = push_stream.cuda_stream
        drawn = [torch.cuda.Event() for _ in range(F)]
    nccl_gather = world > 1 and not fused and not ras_push
    local = torch.empty((F, RAS_H, RAS_W), dtype=torch.int32, device="cuda") if nccl_gather else None
    gathered = torch.empty((n_frames, RAS_H, RAS_W), dtype=torch.int32, device="cuda") if (rank == 0 and nccl_gather) else None
    cams = {k: scenes.lesson_camera(ren, 8, orbit_t(k), RAS_W, RAS_H) for k in range(n_frames * (args.steps + args.warmup + 6))}

    # frames of an animation batch are independent: each of the F targets gets its own CUDA stream, so the short
    # dependent kernel chains of different frames overlap and fill the SMs a single 100k-triangle frame leaves idle
    streams = [torch.cuda.Stream() for _ in range(F)] if args.raster_streams else None
    main_stream = torch.cuda.current_stream()

    def frame(j, k, odd=False):
        raster, g = rasters_b[j] if (odd and in_store) else rasters[j]
        if streams is not None:
            torch.cuda.set_stream(streams[j])          # (the `with torch.cuda.stream()` context costs ~15 us of Python)
        lessons.set_transforms(ren, g, *cams[k])
        lessons.render_frame(ren, raster, vb)
        if ras_push and rank != 0:
            drawn[j].record()
            push_stream.wait_event(drawn[j])
            store.push((n_frames if odd else 0) + my_frames[j], raster.get_render_target().ptr, raster.content_rect, push_ptr)
        if streams is not None:
            torch.cuda.set_stream(main_stream)

    def fork():
        if streams is not None:
            for st in streams:
                st.wait_stream(main_stream)

    def join():
        if streams is not None:
            for st in streams:
                main_stream.wait_stream(st)
        if ras_push:
            main_stream.wait_stream(push_stream)

    def gather():
        if fused or ras_push:
            store.commit()
        elif world > 1:
            for j in range(F):
                local[j].copy_(rasters[j][0].get_render_target().buffer.tensor().view(torch.int32).view(RAS_H, RAS_W))
            parallel.gather_frames(local, gathered, n_frames)

    def step(s):
        fork()
        for j, k in enumerate(my_frames):
            frame(j, s * n_frames + k, odd=bool(s & 1))
        join()
        gather()

    for s in range(args.warmup):
        step(s)
    barrier_sync(world)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        step(args.warmup + s)
    e1.record()
    barrier_sync(world)
    if ras_push:   # untimed check, as for the ray-cast frames: the slots of the last step hold exactly what this rank rendered
        odd_last = bool((args.warmup + args.steps - 1) & 1)
        okf = 1
        if rank != 0:
            for j in range(F):
                slot = store.frame((n_frames if odd_last else 0) + my_frames[j]).view(torch.int32)
                okf &= int(torch.equal(slot, rasters[j][0].get_render_target().buffer.tensor().view(torch.int32).view(-1)))
        flag = torch.tensor([okf], dtype=torch.int32, device="cuda")
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
        assert bool(flag.item()), "raster frames in rank 0's frame store differ from the frames the ranks rendered"
    ms = max_over_ranks(e0.elapsed_time(e1), world)
    tris_total = N_TRIS * n_frames * args.steps
    value = tris_total / (ms * 1e-3) / 1e6
    frame_ms = ms / (args.steps * F)

    host = [torch.zeros((RAS_H, RAS_W), dtype=torch.int32).pin_memory() for _ in range(2)]   # cleared, like the frames
    copy_stream = torch.cuda.Stream()
    copy_ptr = copy_stream.cuda_stream
    done = [torch.cuda.Event() for _ in range(2)]
    reader = parallel.SparseFrameCopier(RAS_W, RAS_H)
    full_rect = (0, 0, RAS_W - 1, RAS_H - 1)

    e2e_rasters = rasters if not fused else [lessons.build_lesson08(ren, ren.create_presenter(RAS_W, RAS_H).get_render_target())
                                             for _ in range(F)]

    def e2e_step(s):
        for j, k in enumerate(my_frames):
            raster, g = e2e_rasters[j]
            cam = scenes.lesson_camera(ren, 8, orbit_t(s * n_frames + k), RAS_W, RAS_H)      # host matrices every frame
            lessons.set_transforms(ren, g, *cam)
            lessons.render_frame(ren, raster, vb)
            done[j % 2].record()
            copy_stream.wait_event(done[j % 2])
            # the frame is the clear colour outside Raster.content_rect (the projected bounding box of what was drawn), and so
            # is the initially cleared host frame outside the content it held before: one pitched D2H copy of the union
            reader.copy(j % 2, host[j % 2].data_ptr(), raster.get_render_target().ptr,
                        raster.content_rect if args.sparse_readback else full_rect, copy_ptr)
        torch.cuda.current_stream().wait_stream(copy_stream)

    e2e_step(args.steps + args.warmup)
    barrier_sync(world)
    k_e2e = max(2, min(args.steps, 5))
    reader.bytes_moved = 0
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for s in range(k_e2e):
        e2e_step(args.steps + args.warmup + 1 + s)
    g1.record()
    barrier_sync(world)
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), world)
    last = e2e_rasters[F - 1][0].get_render_target().buffer.tensor().view(torch.int32).view(RAS_H, RAS_W).cpu()
    e2e_ok = bool(torch.equal(host[(F - 1) % 2], last))
    assert e2e_ok, "sparse read-back: the host frame differs from the device frame"
    hbm_peak, peak_src = peaks()
    alg_bytes = 3 * N_TRIS * 32 + RAS_W * RAS_H * 20                               # SURVEY.md section 8(d), per frame
    achieved = alg_bytes / (frame_ms * 1e-3) / 1e9
    return {
        "metric": METRIC_RAS, "value": value, "unit": "Mtris/s", "ms_per_step": ms / args.steps,
        "config": {"workload": "configs[1]: dragon100k rasterization 1920x1080, lesson08 shaders, clear + clear + draw_triangles per frame",
                   "frames_per_rank_per_step": F, "l2": "8 independent raster targets per rank (~300 MB of key/colour/record buffers > L2)",
                   "streams": "one CUDA stream per frame target (frames of a batch are independent)" if streams else "single stream",
                   **({"gather": "EXPERIMENTAL copy-engine push of Raster.content_rect, verified against the local frames"} if ras_push else {})},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": ncu_traffic().get("raster_frame"), "peak_source": peak_src, "unit_of_work": "one frame = 5 kernels "
                     "(2 clears, raster_kernel, coverage_kernel, resolve_kernel); algorithmic bytes are defined per frame",
                     "frame_ms": frame_ms, "algorithmic_bytes_per_frame": alg_bytes},
        "e2e": {"value": N_TRIS * n_frames * k_e2e / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": 192 * F,
                "d2h_bytes_per_step": reader.bytes_moved // k_e2e, "frame_bytes_per_step": 4 * RAS_W * RAS_H * F,
                "readback_verified": e2e_ok, "steps": k_e2e,
                "note": "per frame: host matrices -> mapped(globals) -> clear, clear, draw_triangles -> async pitched D2H copy into a pinned, "
                        "initially cleared host frame" + (" of Raster.content_rect (projected bounding box of the drawn mesh, united with "
                        "the rect of the frame the host buffer held before)" if args.sparse_readback else " of the whole frame")},
        "gpu_launches": 5 * F * args.steps,
    }


# ---------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------

_CPU_BVH = {}


def cpu_raycast(rows, n_frames, first=0, min_seconds=0.0):
    """CPU BVH closest hit + shade for >= n_frames 4K frames (until min_seconds); returns (Mrays/s, seconds, frames).
    The BVH build is excluded, as on the GPU side."""
    import oracle
    from oracle import host_math as hm
    if id(rows) not in _CPU_BVH:
        _CPU_BVH[id(rows)] = oracle.bvh_build(rows)
    bvh = _CPU_BVH[id(rows)]
    t0 = time.perf_counter()
    k = first
    while k < first + n_frames or time.perf_counter() - t0 < min_seconds:
        W = hm.matmul(hm.scale(1.0), hm.rotate(orbit_t(k), (0, 1, 0)))
        V = hm.look_at((0, 0.3, 2), (0, 0, 0), (0, 1, 0))
        P = hm.perspective(aspect_ratio=RAY_W / RAY_H)
        cam = hm.camera_frame(V, P, W)
        rays = oracle.primary_rays(cam, RAY_W, RAY_H)
        t, ids, u, v = oracle.bvh_raycast(bvh, rays)
        oracle.shade_hits(8, rows, ids, u, v)
        k += 1
    dt = time.perf_counter() - t0
    return RAY_W * RAY_H * (k - first) / dt / 1e6, dt, k - first


def cpu_raster(rows, n_frames, first=0, min_seconds=0.0):
    import oracle
    from oracle import host_math as hm
    t0 = time.perf_counter()
    k = first
    while k < first + n_frames or time.perf_counter() - t0 < min_seconds:
        W = hm.matmul(hm.scale(1.0), hm.rotate(orbit_t(k), (0, 1, 0)))
        V = hm.look_at((0, 0.3, 1.0), (0, 0, 0), (0, 1, 0))
        P = hm.perspective(aspect_ratio=RAS_W / RAS_H)
        oracle.draw_triangles(8, RAS_W, RAS_H, rows, np.concatenate([W.ravel(), V.ravel(), P.ravel()]))
        k += 1
    dt = time.perf_counter() - t0
    return N_TRIS * (k - first) / dt / 1e6, dt, k - first


def reference_arm(args):
    """Times the CPU oracle on this box's host cores, same metric/config as our arm."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    from rendertoy_b200 import scenes
    oracle.build()
    rows = scenes.dragon(N_TRIS)
    cores = oracle.num_threads()
    fn, metric, unit, sample = (cpu_raster, METRIC_RAS, "Mtris/s", "1 frame of 1920x1080 x 100000 triangles per step") \
        if args.path == "raster" else (cpu_raycast, METRIC_RAY, "Mrays/s", "1 frame of 3840x2160 (8.29 Mrays) per step")
    fn(rows, 1, 0)   # builds the CPU BVH (excluded, as on the GPU side) and faults everything in
    for s in range(args.warmup):
        fn(rows, 1, s)
    t0 = time.perf_counter()
    vals = [fn(rows, 1, args.warmup + s)[0] for s in range(args.steps)]
    wall = time.perf_counter() - t0
    units = (RAY_W * RAY_H if args.path != "raster" else N_TRIS) * args.steps
    value = units / wall / 1e6
    kind_note = ("oracle port of the reference pipeline (rendering/_raster.py kernels restated in C + OpenMP; pyopencl is not installable here)"
                 if args.path == "raster" else
                 "CPU BVH of oracle/raycast_oracle.c: the reference has NO ray caster (rendering/_raycaster.py:35-36 is `pass`)")
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": ("configs[1]: dragon100k rasterization 1920x1080, lesson08 shaders" if args.path == "raster" else
                                "configs[3]: dragon100k raycast 3840x2160, lesson06 camera orbit, closest hit + Lambert shade"),
                   "note": kind_note},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "per_step": vals,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="raycast", choices=["raycast", "raster"], help="which half of the metric is the primary line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--raycast-streams", type=int, default=4, help="raycast frames of a step alternate over this many CUDA streams")
    ap.add_argument("--raster-streams", type=int, default=1, help="1: one CUDA stream per raster frame target (default), 0: single stream")
    ap.add_argument("--dense-gather", dest="sparse", action="store_false",
                    help="--gather copy: push whole frames instead of the rect that can differ from the clear colour")
    ap.add_argument("--dense-readback", dest="sparse_readback", action="store_false",
                    help="e2e: read whole ray-cast frames back instead of the rect that can differ from the clear colour")
    ap.add_argument("--view-refit", type=int, default=0, help="EXPERIMENTAL, unmeasured: tightening passes over the screen-space nodes "
                    "(rt_raycast_set_view_refit); 0 = the measured path")
    ap.add_argument("--region-amax", type=float, default=0.0, help="EXPERIMENTAL, unmeasured: two-level region traversal with this frontier "
                    "threshold in tiles (rt_raycast_set_region_traversal); 0 = the measured path")
    ap.add_argument("--raster-gather", default="peer", choices=["peer", "copy"],
                    help="N>1, raster frames: peer = the kernels store into rank 0's frame store (default, measured); copy = EXPERIMENTAL, "
                         "not yet measured: render locally, push Raster.content_rect with the copy engine")
    ap.add_argument("--gather", default="copy", choices=["peer", "copy", "nccl"],
                    help="N>1, raycast frames: copy = ranks render locally and a copy engine pushes each finished frame into rank 0's "
                         "IPC-mapped frame store while the next frame traces (default: fastest from N=4 up); peer = the kernels store "
                         "straight into that frame store over NVLink (fused); nccl = send/recv gather.  Raster frames: peer unless nccl")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        reference_arm(args)
        return

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (rendertoy_b200 has no CPU path); use --impl reference for the CPU oracle")
    # stdout carries exactly ONE line, the JSON: whatever libraries print there while the benchmark runs (NCCL's version
    # banner, for one) is sent to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world, local = dist_setup(args.gpus)
    if args.view_refit or args.region_amax:
        from rendertoy_b200 import _native
        _native.call("rt_raycast_set_view_refit", args.view_refit)
        _native.call("rt_raycast_set_region_traversal", args.region_amax)
    ray, rows = bench_raycast(args, rank, world)
    if args.view_refit or args.region_amax:
        ray["config"]["experimental"] = {"view_refit_passes": args.view_refit, "region_amax_tiles": args.region_amax}
    ras = bench_raster(args, rank, world, rows)
    if rank == 0:
        if not args.no_cpu_baseline:
            import oracle
            oracle.build()
            cores = oracle.num_threads()
            cpu_raycast(rows, 1)
            v, dt, nf = cpu_raycast(rows, 2, min_seconds=10.0)
            ray["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                   "sample": f"{nf} frames of 3840x2160 ({nf * RAY_W * RAY_H / 1e6:.1f} Mrays) in {dt:.1f} s; CPU BVH of "
                                             "oracle/raycast_oracle.c (the reference has no ray caster)"}
            cpu_raster(rows, 1)
            v, dt, nf = cpu_raster(rows, 8, min_seconds=10.0)
            ras["cpu_baseline"] = {"value": v, "unit": "Mtris/s", "cores": cores, "kind": "port",
                                   "sample": f"{nf} frames of 1920x1080 x 100000 triangles in {dt:.1f} s; restated reference pipeline "
                                             "(oracle/raster_oracle.c, OpenMP)"}
        primary, secondary = (ray, ras) if args.path == "raycast" else (ras, ray)
        if args.path == "raster":
            for k in ("n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "clocks"):
                primary.setdefault(k, ray.get(k))
        primary["secondary"] = secondary
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(primary) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
