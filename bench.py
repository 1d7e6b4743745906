#!/usr/bin/env python
"""bench.py -- headline benchmark of rendertoy_b200 (contract: DESIGN.md section 5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--path raycast|raster]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

BASELINE.json's metric has two halves; each is measured on the config it is quoted on, ONE JSON line:

  primary     "Mrays/s closest-hit (dragon, 4K)" -> configs[3]: dragon100k, 3840x2160, lesson06 camera orbit, fused primary
              rays + closest hit + Lambert shade.  Animation batch: frames k = rank (mod N), gathered to rank 0 (weak scaling)
  secondary   "Mtris/s raster"                   -> configs[1]: dragon100k, 1920x1080, lesson08 shaders, clear + clear +
              draw_triangles per frame, same partition
  tiles       configs[3] AS WRITTEN: every 4K frame split over the N ranks by image-space row stripes (parallel.BAND rows,
              stripe s -> rank s % N), stripes gathered into rank 0's frame (strong scaling); same for the raster frame
  config4     configs[4]: the 10M-triangle instanced scene, 256-frame orbit at 1080p, raster + ray cast, frames k = rank (mod N)

A step = RAY_FRAMES (RAS_FRAMES) frames per rank, so that K = 10 steps keep the GPU busy for >= 0.5 s.  `value` is device-timed
whole-job throughput with mesh/BVH resident; `e2e` is the same work driven through the public `rendering` API with host-side
inputs and every frame read back to pinned host memory.  `--impl reference` times the CPU oracle (oracle/: the restated reference
raster pipeline; the reference has no ray caster, so that half is our CPU BVH definition) on all host cores.
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

N_TRIS = int(os.environ.get("RENDERTOY_B200_TRIS", "100000"))   # (the override is a dev knob: host-bound or GPU-bound?)
RAY_W, RAY_H = 3840, 2160
RAS_W, RAS_H = 1920, 1080
ORBIT = 256       # World = rotate(2*pi*k/256, y)  
SUB = int(os.environ.get("RENDERTOY_B200_SUB", "8"))   # frames per sub-batch = distinct frame targets a rank cycles through (8 x 33 MB > L2)
RAY_FRAMES = int(os.environ.get("RENDERTOY_B200_RAY_FRAMES", "512"))     # frames per rank and step
RAS_FRAMES = int(os.environ.get("RENDERTOY_B200_RAS_FRAMES", "1024"))
TILE_FRAMES = int(os.environ.get("RENDERTOY_B200_TILE_FRAMES", "256"))   # frames per step (all ranks together) in the tile partition

METRIC_RAY = "Mrays/s closest-hit (dragon, 4K)"
METRIC_RAS = "Mtris/s raster"

# `config` names the workload and nothing else: both arms (ours, --impl reference) print exactly these dicts
CONFIG_RAY = {"workload": "configs[3]: dragon100k (synthetic stand-in for the missing dragon.obj, 100000 triangles) raycast 3840x2160, "
                          "lesson06 camera orbit, primary rays + closest hit + Lambert shade",
              "triangles": N_TRIS, "width": RAY_W, "height": RAY_H,
              "camera": "tutorials/lesson06_loading_obj.py:74-83, World = rotate(2*pi*k/256, y), frame k of a 256-frame orbit"}
CONFIG_RAS = {"workload": "configs[1]: dragon100k rasterization 1920x1080, lesson08 shaders, clear + clear + draw_triangles per frame",
              "triangles": N_TRIS, "width": RAS_W, "height": RAS_H,
              "camera": "tutorials/lesson08_rasterization.py:90-99, World = rotate(2*pi*k/256, y), frame k of a 256-frame orbit"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_fact(name):
    v = ncu_facts().get(name)
    if isinstance(v, dict):
        return v
    return {} if v is None else {"dram_bytes_per_launch": v}


def ncu_facts():
    """Per-launch facts read off the committed ncu --set full captures (profiles/traffic.json): DRAM bytes, executed warp
    instructions, issue-slot utilisation.  bench.py never runs under a profiler; these are static properties of the same
    kernels on the same scene, used to state the binding roofline (instruction issue) next to the HBM one."""
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
            time.sleep(0.02)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# scene / distributed helpers
# ---------------------------------------------------------------------------------------------------------

def orbit_t(k):
    return 2.0 * math.pi * (k % ORBIT) / ORBIT


def upload(ren, rows):
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    return vb


_CAMS = {}


def ray_camera(ren, k, lesson=6, w=RAY_W, h=RAY_H):
    """12-float camera frame of orbit frame k (cached: the orbit has 256 distinct frames)."""
    key = (lesson, k % ORBIT, w, h)
    if key not in _CAMS:
        from rendering._raycaster import camera_frame
        from rendertoy_b200 import scenes
        world, view, proj = scenes.lesson_camera(ren, lesson, orbit_t(k), w, h)
        _CAMS[key] = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
    return _CAMS[key]


def raster_camera(ren, k, w=RAS_W, h=RAS_H):
    """(World, View, Proj) of orbit frame k as the (c_float * 48) Raster.draw_frame takes (cached per orbit position: the
    device-resident loop's inputs are prepared before the timed region; the e2e loop computes its matrices every frame)."""
    key = ("ras", k % ORBIT, w, h)
    if key not in _CAMS:
        import ctypes
        from rendertoy_b200 import scenes
        from rendering._raster import transforms48
        _CAMS[key] = (ctypes.c_float * 48)(*transforms48(*scenes.lesson_camera(ren, 8, orbit_t(k), w, h)).tolist())
    return _CAMS[key]


def dist_setup():
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


def all_ranks(v, world):
    """every rank's value, in rank order"""
    import torch
    import torch.distributed as dist
    if world > 1:
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[dist.get_rank()] = v
        dist.all_reduce(t)
        return [float(x) for x in t.cpu()]
    return [float(v)]


def max_over_ranks(ms, world):
    return max(all_ranks(ms, world))


def all_ok(ok, world):
    import torch
    if world == 1:
        return bool(ok)
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
    torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
    return bool(flag.item())


class Streams:
    """Independent frames alternate over a few CUDA streams so one frame's tail overlaps its neighbours; forked from / joined
    into the main stream only where something has to be ordered (timing events, the commit of a gather)."""

    def __init__(self, n):
        import torch
        self.torch = torch
        self.main = torch.cuda.current_stream()
        self.streams = [torch.cuda.Stream() for _ in range(n)] if n > 1 else []

    def fork(self):
        for st in self.streams:
            st.wait_stream(self.main)

    def use(self, i):
        if self.streams:
            st = self.streams[i % len(self.streams)]
            self.torch.cuda.set_stream(st)       # (the `with torch.cuda.stream()` context costs ~15 us of Python)
            return st
        return self.main

    def join(self, *extra):
        if self.streams:
            self.torch.cuda.set_stream(self.main)
            for st in self.streams:
                self.main.wait_stream(st)
        for st in extra:
            self.main.wait_stream(st)


# ---------------------------------------------------------------------------------------------------------
# animation batch (frames k = rank mod N), both paths: the frame loop with its gather
# ---------------------------------------------------------------------------------------------------------

class FrameLoop:
    """Runs `frames` frames per rank and step through `render(i, k, target) -> content rect` (i: frame of the step, k: global
    frame number = orbit position, target: what make_target() built), cycling SUB local targets, and gathers the frames into
    rank 0's frame store:

      gather == "copy": ranks != 0 render locally, a copy engine pushes each finished frame's content rect into its slot
                        (cover_rect: only what can differ from what the slot holds); rank 0 renders in place
      gather == "peer": every rank's kernels store straight into the slots over NVLink (the render targets ARE the slots)
      gather == "nccl": render locally, grouped send/recv per sub-batch (baseline)

    The store is a ring of 2 * commit_every sub-batches (SUB frames per rank each); a stream-ordered 4-byte all-reduce every
    `commit_every` sub-batches tells rank 0 those frames are complete, and a slot is rewritten only two commits later."""

    def __init__(self, ren, world, rank, W, H, frames, gather, commit_every, n_streams, make_target, tile_push=False):
        import torch
        from rendertoy_b200 import parallel
        self.torch, self.ren, self.world, self.rank, self.W, self.H, self.frames = torch, ren, world, rank, W, H, frames
        self.parallel = parallel
        self.C = max(1, commit_every)
        self.ring = 2 * self.C                          # sub-batches in the ring
        assert frames % (SUB * self.C) == 0, "frames per step must be a multiple of SUB * commit_every"
        self.store = None
        self.gather = gather if world > 1 else "none"
        if self.gather in ("copy", "peer"):
            self.store = parallel.FrameStore(self.ring * SUB * world, W, H)
            if not self.store.ok:
                self.gather = "nccl"
        self.in_store = self.gather == "peer" or (self.gather == "copy" and rank == 0)
        if self.in_store:      # one target object per slot this rank renders into
            self.targets = [make_target(self.store.frame(self.slot(q, j))) for q in range(self.ring) for j in range(SUB)]
        else:
            self.targets = [make_target(None) for _ in range(SUB)]
        self.streams = Streams(n_streams)
        self.push_stream = torch.cuda.Stream() if (self.gather == "copy" and rank != 0) else None
        self.pushed = [None] * SUB                      # event: the push out of local target j has finished reading it
        self.rendered_ev = [torch.cuda.Event() for _ in range(SUB)]
        self.pushed_ev = [torch.cuda.Event() for _ in range(SUB)]
        self.pushed_bytes = 0
        self.tile_push = tile_push      # gather == "copy": rt_push_tiles (a kernel, non-clear 32x32 tiles) instead of rt_copy_rect (copy engine, content rect)
        self.q = 0                                      # global sub-batch counter
        if self.gather == "nccl":
            self.local = torch.empty((SUB, H, W), dtype=torch.int32, device="cuda")
            self.gathered = torch.empty((SUB * world, H, W), dtype=torch.int32, device="cuda") if rank == 0 else None

    def slot(self, q, j):
        return ((q % self.ring) * SUB + j) * self.world + self.rank

    def target(self, q, j):
        return self.targets[(q % self.ring) * SUB + j] if self.in_store else self.targets[j]

    def step(self, s, render, sparse=True):
        torch = self.torch
        full = (0, 0, self.W - 1, self.H - 1)
        self.streams.fork()
        for b in range(self.frames // SUB):
            q = self.q
            for j in range(SUB):
                i = b * SUB + j
                k = (s * self.frames + i) * self.world + self.rank          # global frame number -> orbit position
                st = self.streams.use(i)
                if self.push_stream is not None and self.pushed[j] is not None:
                    st.wait_event(self.pushed[j])                              # the target's previous frame has left
                tgt = self.target(q, j)
                content = render(i, k, tgt)
                if self.push_stream is not None:
                    self.rendered_ev[j].record(st)
                    self.push_stream.wait_event(self.rendered_ev[j])
                    if self.tile_push:
                        self.store.push_tiles(self.slot(q, j), tgt_ptr(tgt), self.push_stream.cuda_stream)
                    else:
                        self.pushed_bytes += self.store.push(self.slot(q, j), tgt_ptr(tgt), content if sparse else full, self.push_stream.cuda_stream)
                    self.pushed_ev[j].record(self.push_stream)
                    self.pushed[j] = self.pushed_ev[j]
            self.q += 1
            if self.gather == "nccl":
                self.streams.join()
                for j in range(SUB):
                    self.local[j].copy_(tgt_tensor(self.targets[j]).view(torch.int32).view(self.H, self.W))
                self.parallel.gather_frames(self.local, self.gathered, SUB * self.world)
                self.streams.fork()
            elif self.store is not None and self.q % self.C == 0:
                self.streams.join(*([self.push_stream] if self.push_stream is not None else []))
                self.store.commit()
                self.streams.fork()
        self.streams.join(*([self.push_stream] if self.push_stream is not None else []))

    def close(self):
        if self.store is not None:
            self.torch.cuda.synchronize()
            barrier_sync(self.world)
            self.targets = None
            self.store.close()
            self.store = None

    def verify_last(self):
        """Untimed: the slots of the last sub-batch hold exactly what this rank rendered locally (read back over NVLink)."""
        if self.gather != "copy":
            return None
        ok = True
        if self.rank != 0:
            q = self.q - 1
            for j in range(SUB):
                slot = self.store.frame(self.slot(q, j)).view(self.torch.int32)
                ok &= bool(self.torch.equal(slot, tgt_tensor(self.targets[j]).view(self.torch.int32).view(-1)))
        return all_ok(ok, self.world)


def tgt_ptr(t):
    """device address of a frame target: an Image, or a (Raster, globals...) tuple"""
    return (t[0].get_render_target() if isinstance(t, tuple) else t).ptr


def tgt_tensor(t):
    return (t[0].get_render_target() if isinstance(t, tuple) else t).buffer.tensor()


def gather_text(loop, sparse):
    if loop.world == 1:
        return "single GPU, no gather"
    if loop.gather == "copy" and loop.tile_push:
        return ("ranks != 0 render locally and a small kernel on the producing GPU (rt_push_tiles) stores the 32x32-pixel tiles of each finished frame "
                "that hold something other than the clear colour -- or did the last time the slot was written -- straight into rank 0's frame store "
                "(CUDA IPC peer memory, cleared at start) while the next frames render; rank 0 renders in place; "
                f"frame store = ring of {loop.ring} sub-batches x {SUB} frames per rank, one stream-ordered 4-byte all-reduce per {loop.C} sub-batches")
    ring = f"frame store = ring of {loop.ring} sub-batches x {SUB} frames per rank, one stream-ordered 4-byte all-reduce per {loop.C} sub-batches"
    if loop.gather == "peer":
        return "every rank's kernel stores its pixels straight into rank 0's frame store over NVLink (CUDA IPC peer memory); " + ring
    if loop.gather == "copy":
        return ("ranks != 0 render locally and push each finished frame into rank 0's frame store (CUDA IPC peer memory, cleared at start) with an "
                "async pitched copy-engine transfer that overlaps the next frames" + ("; only the pixel rect that can differ from the slot's content "
                "travels (union of this frame's and the previous frame's content rect: the frame is the clear colour elsewhere)" if sparse else "")
                + "; rank 0 renders in place; " + ring)
    return "framebuffers gathered to rank 0 with NCCL send/recv per sub-batch"


# ---------------------------------------------------------------------------------------------------------
# primary: ray casting
# ---------------------------------------------------------------------------------------------------------

def bench_raycast(args, rank, world, rows, vb, W=RAY_W, H=RAY_H, lesson=6, frames=None, full=True, n_tris=N_TRIS):
    """full: also e2e, the isolated-launch figure, the instrumented pass and the frame-filling camera (the primary line);
    otherwise only the device-timed loop (config4)."""
    import torch
    import rendering as ren
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import parallel

    F = frames or RAY_FRAMES
    rc = Raycaster([ren.Mesh(vb, None)])
    loop = FrameLoop(ren, world, rank, W, H, F, args.gather, args.commit_every, args.raycast_streams,
                     lambda mem: ren.Image(W, H, ren._core.RGBA, memory=mem) if mem is not None else ren.create_image2d(W, H, ren._core.RGBA))
    for k in range(ORBIT):
        ray_camera(ren, k, lesson, W, H)

    def render(i, k, tgt):
        return rc.render(tgt, ray_camera(ren, k, lesson, W, H))

    for s in range(args.warmup):
        loop.step(s, render, args.sparse)
    sampler = ClockSampler(torch.cuda.current_device()); sampler.start()
    barrier_sync(world)
    loop.pushed_bytes = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        loop.step(args.warmup + s, render, args.sparse)
    e1.record()
    barrier_sync(world)
    clocks = sampler.result()
    gather_ok = loop.verify_last()
    assert gather_ok is not False, "frames in rank 0's frame store differ from the frames the ranks rendered"
    ms_local = e0.elapsed_time(e1)
    ms_ranks = all_ranks(ms_local, world)
    ms = max(ms_ranks)
    push_bytes_step = all_ranks(loop.pushed_bytes / max(args.steps, 1), world)
    value = W * H * F * world * args.steps / (ms * 1e-3) / 1e6
    kernel_ms = ms_local / (args.steps * F)      # timed region / launch pairs in it (frames overlap on the streams)
    out = {"value": value, "ms": ms, "ms_ranks": ms_ranks, "kernel_ms": kernel_ms, "clocks": clocks, "frames": F,
           "gather": gather_text(loop, args.sparse), "gather_verified": gather_ok, "push_bytes_step": push_bytes_step,
           "launches": (2 + (1 if args.view_refit else 0)) * F * args.steps, "view_nodes": rc.n_triangles <= 1 << 18}
    if not full:
        loop.close()
        return out

    targets = loop.targets[:SUB]
    # the same launch pair alone on the GPU
    iso = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(SUB)]
    for j in range(SUB):
        iso[j][0].record(); rc.render(targets[j], ray_camera(ren, j * world + rank)); iso[j][1].record()
    torch.cuda.synchronize()
    out["kernel_ms_alone"] = float(np.mean([a.elapsed_time(b) for a, b in iso]))

    # instrumented pass: node visits / triangle tests per ray
    stats = torch.zeros(3, dtype=torch.int64, device="cuda")
    rc.render(targets[0], ray_camera(ren, rank), stats=stats)
    torch.cuda.synchronize()
    out["stats"] = tuple(int(x) for x in stats.cpu())

    # the frame-filling camera (lesson08's, eye at distance 1: the mesh covers ~38 % of the 4K frame, the traced rect ~90 %)
    streams = Streams(args.raycast_streams)
    n_ff = 64

    def ff_pass():
        streams.fork()
        for i in range(n_ff):
            streams.use(i)
            rc.render(targets[i % SUB], ray_camera(ren, i * world + rank, 8))
        streams.join()
    ff_pass()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record(); ff_pass(); ff_pass(); f1.record()
    torch.cuda.synchronize()
    ff_ms = f0.elapsed_time(f1) / (2 * n_ff)
    stats.zero_()
    rc.render(targets[0], ray_camera(ren, rank, 8), stats=stats)
    torch.cuda.synchronize()
    ff_stats = tuple(int(x) for x in stats.cpu())
    out["frame_filling"] = {"camera": "tutorials/lesson08_rasterization.py:90-99 (eye (0,0.3,1)) at 3840x2160, same orbit", "value": W * H / ff_ms / 1e3,
                            "unit": "Mrays/s", "ms_per_frame": ff_ms, "rays_traced_fraction": ff_stats[2] / (W * H), "n_gpus": 1,
                            "inner_node_visits_per_ray": ff_stats[0] / (W * H), "triangle_tests_per_ray": ff_stats[1] / (W * H),
                            "note": "per GPU, no gather; frames alternate over the same streams"}

    # ---- e2e: public API, host inputs, every frame read back to pinned host memory
    host = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]   # cleared, like the frames
    copy_stream = torch.cuda.Stream()
    copy_ptr = copy_stream.cuda_stream
    from rendertoy_b200 import scenes
    from rendering._raycaster import camera_frame
    # e2e delivers frames to HOST memory: every rank reads its own frames back over its own PCIe link, so the GPU-side
    # gather is not on this path (local targets, no collective)
    e2e_targets = [ren.create_image2d(W, H, ren._core.RGBA) for _ in range(SUB)] if (loop.in_store and world > 1) else targets
    e2e_streams = Streams(2 if args.raycast_streams > 1 else 1)       # measured: 43.6 Grays/s with 2, 40.3 with 1, 38.8 with 4
    full_rect = (0, 0, W - 1, H - 1)

    rendered = [torch.cuda.Event() for _ in range(SUB)]
    read_done = [torch.cuda.Event() for _ in range(SUB)]

    def e2e_pass(n, s0, reader, sparse):
        """n frames: host matrices -> camera frame -> render -> async D2H copy into one of two pinned host frames"""
        e2e_streams.fork()
        for i in range(n):
            k = (s0 + i) * world + rank
            j = i % SUB
            world_m, view, proj = scenes.lesson_camera(ren, lesson, orbit_t(k), W, H)     # host inputs, computed every frame
            cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world_m, dtype=ren.float4x4))
            st = e2e_streams.use(i)
            if i >= SUB:
                st.wait_event(read_done[j])      # the frame this target held has been read back
            content = rc.render(e2e_targets[j], cam)
            rendered[j].record(st)
            copy_stream.wait_event(rendered[j])
            if sparse == "tiles":     # a kernel stores the non-clear 32x32 tiles straight into the pinned host frame (unified addressing)
                reader.copy(i % 2, host[i % 2].data_ptr(), e2e_targets[j].ptr, copy_ptr)
            else:
                reader.copy(i % 2, host[i % 2].data_ptr(), e2e_targets[j].ptr, content if sparse else full_rect, copy_ptr)
            read_done[j].record(copy_stream)
        e2e_streams.join(copy_stream)

    def e2e_measure(n_frames, sparse):
        tiles = sparse == "tiles"
        reader = parallel.TileFrameCopier(W, H) if tiles else parallel.SparseFrameCopier(W, H)
        for hbuf in host:
            hbuf.zero_()
        e2e_pass(2 * SUB, 0, reader, sparse)
        barrier_sync(world)
        if tiles:
            reader.bytes.zero_()
            torch.cuda.synchronize()
        else:
            reader.bytes_moved = 0
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        e2e_pass(n_frames, 2 * SUB, reader, sparse)
        g1.record()
        barrier_sync(world)
        ms_e = max_over_ranks(g0.elapsed_time(g1), world)
        j_last = (n_frames - 1) % SUB
        ok = bool(torch.equal(host[(n_frames - 1) % 2], e2e_targets[j_last].buffer.tensor().view(torch.int32).view(H, W).cpu()))
        assert ok, "read-back: the host frame differs from the device frame"
        return W * H * n_frames * world / (ms_e * 1e-3) / 1e6, (reader.bytes_moved() if tiles else reader.bytes_moved), ok, ms_e

    k_e2e = max(2, min(args.steps, 5))
    n_e2e = k_e2e * F
    mode = ("tiles" if args.readback == "tiles" else True) if args.sparse_readback else False
    e2e_value, e2e_bytes, e2e_ok, e2e_ms = e2e_measure(n_e2e, mode)
    dense_value, dense_bytes, _, dense_ms = e2e_measure(max(SUB * 8, n_e2e // 8), False)
    out["e2e"] = {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": 192 * F, "d2h_bytes_per_step": e2e_bytes // k_e2e,
                  "frame_bytes_per_step": 4 * W * H * F, "readback_verified": e2e_ok, "steps": k_e2e, "timed_region_ms": e2e_ms,
                  "dense_readback": {"value": dense_value, "unit": "Mrays/s", "d2h_bytes_per_frame": 4 * W * H, "timed_region_ms": dense_ms,
                                     "note": "the same loop reading every frame back whole (33 MB per frame over PCIe)"},
                  "readback": "tiles" if mode == "tiles" else ("rect" if mode else "dense"),
                  "note": "per frame: host matrices -> camera frame -> rt_raycast_primary -> " + (
                          "rt_push_tiles: a kernel stores the 32x32-pixel tiles that are not the clear colour (or were not in the frame the host "
                          "buffer held before) straight into the pinned, initially cleared host frame over PCIe (unified addressing)" if mode == "tiles" else
                          "async pitched D2H copy into a pinned, initially cleared host frame" + (" of the pixel rect that can differ from the clear "
                          "colour (union of the scene's projected bounds of this frame and of the frame the host buffer held before)" if mode
                          else " of the whole 33 MB frame")) + "; the host frame is complete and checked against the device frame after the timed "
                          "region; every rank reads back its own frames"}
    del targets, e2e_targets
    loop.close()
    return out


# ---------------------------------------------------------------------------------------------------------
# secondary: rasterization
# ---------------------------------------------------------------------------------------------------------

def bench_raster(args, rank, world, rows, vb, W=RAS_W, H=RAS_H, frames=None, full=True, n_tris=N_TRIS):
    import torch
    import rendering as ren
    from rendertoy_b200 import lessons, parallel

    F = frames or RAS_FRAMES

    def make_target(mem):
        img = ren.Image(W, H, ren._core.RGBA, memory=mem) if mem is not None else ren.create_presenter(W, H).get_render_target()
        return lessons.build_lesson08(ren, img)          # (raster, globals)

    loop = FrameLoop(ren, world, rank, W, H, F, args.raster_gather, args.commit_every, SUB if args.raster_streams else 1, make_target,
                     tile_push=args.raster_push == "tiles")
    for k in range(ORBIT):
        raster_camera(ren, k, W, H)

    def render(i, k, tgt):
        tgt[0].draw_frame(vb, None, raster_camera(ren, k, W, H))      # = set World/View/Proj + clear + clear + draw_triangles
        return tgt[0].content_rect if (loop.push_stream is not None and not loop.tile_push) else None

    for s in range(args.warmup):
        loop.step(s, render, args.sparse)
    barrier_sync(world)
    loop.pushed_bytes = 0
    if loop.store is not None and loop.store.tile_bytes is not None:
        loop.store.tile_bytes.zero_()
        torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        loop.step(args.warmup + s, render, args.sparse)
    e1.record()
    barrier_sync(world)
    gather_ok = loop.verify_last()
    assert gather_ok is not False, "raster frames in rank 0's frame store differ from the frames the ranks rendered"
    ms_ranks = all_ranks(e0.elapsed_time(e1), world)
    ms = max(ms_ranks)
    value = n_tris * F * world * args.steps / (ms * 1e-3) / 1e6
    frame_ms = e0.elapsed_time(e1) / (args.steps * F)
    if loop.store is not None and loop.store.tile_bytes is not None:
        loop.pushed_bytes = int(loop.store.tile_bytes.item())
    push_bytes_step = all_ranks(loop.pushed_bytes / max(args.steps, 1), world)
    out = {"value": value, "ms": ms, "ms_ranks": ms_ranks, "frame_ms": frame_ms, "frames": F, "gather": gather_text(loop, args.sparse),
           "gather_verified": gather_ok, "launches": (4 + (1 if (loop.tile_push and loop.push_stream is not None) else 0)) * F * args.steps,
           "streams": max(1, len(loop.streams.streams)), "push_bytes_step": push_bytes_step}
    if not full:
        loop.close()
        return out

    # one frame alone on the GPU
    raster0, g0_ = loop.targets[0]
    iso = []
    for j in range(8):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for r in range(4):          # four frames back to back on one stream (the first one hides the launch latency of the rest)
            raster0.draw_frame(vb, None, raster_camera(ren, 4 * j + r, W, H))
        b.record()
        iso.append((a, b))
    torch.cuda.synchronize()
    out["frame_ms_alone"] = float(np.mean([a.elapsed_time(b) for a, b in iso])) / 4

    # ---- e2e
    from rendertoy_b200 import scenes
    host = [torch.zeros((H, W), dtype=torch.int32).pin_memory() for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    copy_ptr = copy_stream.cuda_stream
    full_rect = (0, 0, W - 1, H - 1)
    e2e_rasters = [make_target(None) for _ in range(SUB)] if (loop.in_store and world > 1) else loop.targets[:SUB]

    rendered = [torch.cuda.Event() for _ in range(SUB)]
    read_done = [torch.cuda.Event() for _ in range(SUB)]

    def e2e_pass(n, s0, reader):
        main = torch.cuda.current_stream()
        for i in range(n):
            k = (s0 + i) * world + rank
            j = i % SUB
            raster, g = e2e_rasters[j]
            cam = scenes.lesson_camera(ren, 8, orbit_t(k), W, H)      # host matrices every frame
            lessons.set_transforms(ren, g, *cam)
            if i >= SUB:
                main.wait_event(read_done[j])    # the frame this target held has been read back
            lessons.render_frame(ren, raster, vb)
            rendered[j].record(main)
            copy_stream.wait_event(rendered[j])
            if tiles_rb:
                reader.copy(i % 2, host[i % 2].data_ptr(), raster.get_render_target().ptr, copy_ptr)
            else:
                reader.copy(i % 2, host[i % 2].data_ptr(), raster.get_render_target().ptr,
                            raster.content_rect if args.sparse_readback else full_rect, copy_ptr)
            read_done[j].record(copy_stream)
        main.wait_stream(copy_stream)

    tiles_rb = args.sparse_readback and args.readback == "tiles"
    reader = parallel.TileFrameCopier(W, H) if tiles_rb else parallel.SparseFrameCopier(W, H)
    e2e_pass(2 * SUB, 0, reader)
    barrier_sync(world)
    k_e2e = max(2, min(args.steps, 5))
    n_e2e = k_e2e * F
    if tiles_rb:
        reader.bytes.zero_()
        torch.cuda.synchronize()
    else:
        reader.bytes_moved = 0
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    e2e_pass(n_e2e, 2 * SUB, reader)
    g1.record()
    barrier_sync(world)
    e2e_ms = max_over_ranks(g0.elapsed_time(g1), world)
    last = e2e_rasters[(n_e2e - 1) % SUB][0].get_render_target().buffer.tensor().view(torch.int32).view(H, W).cpu()
    e2e_ok = bool(torch.equal(host[(n_e2e - 1) % 2], last))
    assert e2e_ok, "read-back: the host frame differs from the device frame"
    out["e2e"] = {"value": n_tris * n_e2e * world / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": 192 * F,
                  "d2h_bytes_per_step": (reader.bytes_moved() if tiles_rb else reader.bytes_moved) // k_e2e, "frame_bytes_per_step": 4 * W * H * F,
                  "readback_verified": e2e_ok, "steps": k_e2e, "timed_region_ms": e2e_ms,
                  "readback": "tiles" if tiles_rb else ("rect" if args.sparse_readback else "dense"),
                  "note": "per frame: host matrices -> mapped(globals) -> clear, clear, draw_triangles -> " + (
                          "rt_push_tiles: a kernel stores the non-clear 32x32-pixel tiles (and those that were non-clear in the frame the host buffer "
                          "held before) straight into the pinned, initially cleared host frame over PCIe" if tiles_rb else
                          "async pitched D2H copy into a pinned, initially cleared host frame" + (" of Raster.content_rect (projected bounding box of "
                          "the drawn mesh, united with the rect of the frame the host buffer held before)" if args.sparse_readback else " of the whole frame"))}
    del e2e_rasters, raster0, g0_
    loop.close()
    return out


# ---------------------------------------------------------------------------------------------------------
# configs[3] as written: ONE frame split over the ranks by image-space stripes (strong scaling), both paths
# ---------------------------------------------------------------------------------------------------------

def bench_tiles(args, rank, world, rows, vb):
    """Every frame is rendered by ALL ranks: rank r owns the row stripes s = r (mod N) of parallel.BAND rows and renders them with one
    launch (ray cast: rt_raycast_primary(stripes); raster: a scissored draw of the whole mesh) -- into rank 0's frame directly
    (peer stores) or locally followed by ONE 3-D copy-engine push of its stripes.  value = W*H*frames/t resp. T*frames/t."""
    import torch
    import rendering as ren
    from rendering._raycaster import Raycaster
    from rendertoy_b200 import lessons, parallel, _native

    C = max(1, args.commit_every)
    ring = 2 * C * SUB                                  # frames in the ring
    FT = TILE_FRAMES
    assert FT % (C * SUB) == 0
    res = {}
    for path in ("raycast", "raster"):
        W, H = (RAY_W, RAY_H) if path == "raycast" else (RAS_W, RAS_H)
        gather = "none" if world == 1 else (args.tiles_gather if args.tiles_gather != "auto" else ("copy" if path == "raycast" else "peer"))
        store = parallel.FrameStore(ring, W, H) if world > 1 else None
        if store is not None and not store.ok:      # (store.ok is agreed on by all ranks)
            res[path] = {"unavailable": "the tile partition gathers through the CUDA-IPC frame store, which could not be mapped on this box"}
            continue
        stripes = parallel.stripes_of(rank, world) if world > 1 else None
        in_store = store is not None and (gather == "peer" or rank == 0)
        push_stream = torch.cuda.Stream() if (gather == "copy" and rank != 0) else None
        n_targets = ring if in_store else SUB

        def image(i):
            return ren.Image(W, H, ren._core.RGBA, memory=store.frame(i)) if in_store else ren.create_image2d(W, H, ren._core.RGBA)
        if path == "raycast":
            rc = Raycaster([ren.Mesh(vb, None)])
            # every rank repeats the projection + tightening of the whole BVH for every frame: from 4 ranks up two iterations pay
            # better than four, and more streams hide more of those single-wave kernels (measured at N = 8 with graph replay:
            # 4 streams / 4 iterations 255 Grays/s, 8 / 4 -> 301, 8 / 2 -> 318, 2 / 4 -> 168)
            tiles_refit = args.tiles_view_refit if args.tiles_view_refit is not None else (2 if world >= 4 else args.view_refit)
            _native.call("rt_raycast_set_view_refit", tiles_refit)
            targets = [image(i) for i in range(n_targets)]
            streams = Streams(max(args.raycast_streams, SUB) if world > 1 else args.raycast_streams)

            def render(f, tgt):
                return rc.render(tgt, ray_camera(ren, f), stripes=stripes)
        else:
            targets = []
            for i in range(n_targets):
                raster, g = lessons.build_lesson08(ren, image(i))
                if stripes is not None:
                    raster.set_scissor(stripes=stripes)
                targets.append((raster, g))
            streams = Streams(SUB if args.raster_streams else 1)

            def render(f, tgt):
                tgt[0].draw_frame(vb, None, raster_camera(ren, f))
                return tgt[0].content_rect if push_stream is not None else None
        L = C * SUB                                     # frames between two commits
        assert FT % L == 0 and FT % ring == 0 and FT % ORBIT == 0, "a step is whole orbits, whole commit intervals, whole rings"
        rendered_ev = [torch.cuda.Event() for _ in range(SUB)]
        pushed_ev = [torch.cuda.Event() for _ in range(SUB)]

        def interval(f0):
            """enqueue frames f0 .. f0 + L - 1 of the orbit (slot f % ring, local target f % SUB); self-contained: forks from and
            joins into the current main stream, waits only on events recorded inside itself (so it can be captured as a graph)"""
            pushed = [False] * SUB
            streams.fork()
            for f in range(f0, f0 + L):
                st = streams.use(f)
                j = f % SUB
                if push_stream is not None and pushed[j]:
                    st.wait_event(pushed_ev[j])                # the local target's previous frame has left
                tgt = targets[f % ring] if in_store else targets[j]
                content = render(f, tgt)
                if push_stream is not None:
                    rendered_ev[j].record(st)
                    push_stream.wait_event(rendered_ev[j])
                    store.push_stripes(f % ring, tgt_ptr(tgt), content, stripes, push_stream.cuda_stream)
                    pushed_ev[j].record(push_stream)
                    pushed[j] = True
            streams.join(*([push_stream] if push_stream is not None else []))

        graphs = None

        def step():
            for g in range(FT // L):
                if graphs is not None:
                    graphs[g % len(graphs)].replay()
                else:
                    interval(g * L)
                if store is not None:
                    store.commit()

        for _ in range(2):                                  # eager: allocations, and the slots' content history reaches its steady state
            step()
        barrier_sync(world)
        graph_note = "eager launches from Python"
        if args.tiles_graphs:
            # The orbit is a fixed launch sequence (256 cameras, ring slots f % ring): capture each commit interval ONCE as a CUDA
            # graph -- kernels with their baked camera / matrices, the copy-engine pushes, the stream forks and joins -- and replay.
            # The per-frame host cost (Python + ~4 driver calls, ~40 us: the plateau of the eager loop from N = 4 on) disappears.
            try:
                main_stream = streams.main
                cap = torch.cuda.Stream()
                caught = []
                for g in range(ORBIT // L):
                    cg = torch.cuda.CUDAGraph()
                    cap.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.graph(cg, stream=cap):
                        streams.main = torch.cuda.current_stream()
                        interval(g * L)
                    streams.main = main_stream
                    caught.append(cg)
                graphs = caught
                graph_note = f"CUDA-graph replay: the orbit's {ORBIT // L} commit intervals of {L} frames captured once (kernels, pushes, stream forks/joins), one cudaGraphLaunch each"
            except Exception as e:      # e.g. a driver that refuses peer copies inside a capture: stay eager, and say so
                streams.main = main_stream
                torch.cuda.set_stream(main_stream)
                graphs = None
                graph_note = f"eager launches from Python (graph capture failed: {type(e).__name__}: {str(e)[:120]})"
        captured = all_ok(graphs is not None, world)
        if not captured:
            graphs = None
        for _ in range(max(1, min(args.warmup, 3))):
            step()
        barrier_sync(world)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step()
        e1.record()
        barrier_sync(world)
        counter = [FT]
        ms_ranks = all_ranks(e0.elapsed_time(e1), world)
        ms = max(ms_ranks)
        # untimed check on rank 0: the gathered frame of the last step equals a whole-frame render of the same camera
        ok = True
        if world > 1:
            f_last = counter[0] - 1
            if rank == 0:
                if path == "raycast":
                    ref = ren.create_image2d(W, H, ren._core.RGBA)
                    rc.render(ref, ray_camera(ren, f_last))
                    ref_t = ref.buffer.tensor()
                else:
                    raster, g = lessons.build_lesson08(ren, ren.create_image2d(W, H, ren._core.RGBA))
                    raster.draw_frame(vb, None, raster_camera(ren, f_last))
                    ref_t = raster.get_render_target().buffer.tensor()
                torch.cuda.synchronize()
                ok = bool(torch.equal(store.frame(f_last % ring), ref_t))
            ok = all_ok(ok, world)
            assert ok, f"tile partition ({path}): the gathered frame differs from the whole-frame render"
        units = (W * H) if path == "raycast" else N_TRIS
        res[path] = {"metric": METRIC_RAY if path == "raycast" else METRIC_RAS, "unit": "Mrays/s" if path == "raycast" else "Mtris/s",
                     "value": units * FT * args.steps / (ms * 1e-3) / 1e6, "scaling": "strong", "n_gpus": world,
                     "ms_per_frame": ms / (FT * args.steps), "frames_per_step": FT, "timed_region_ms_per_rank": ms_ranks,
                     "gathered_frame_verified": ok if world > 1 else None, "launch": graph_note,
                     "streams": max(1, len(streams.streams)), **({"view_refit_iterations": tiles_refit} if path == "raycast" else {}),
                     "partition": ("single GPU: the whole frame, no stripes" if world == 1 else
                                   f"every frame split into row stripes of {parallel.BAND} rows, stripe s -> rank s % {world}; one launch per rank and frame; "
                                   + ("kernels store their stripes straight into rank 0's frame (peer memory)" if gather == "peer" else
                                      "ranks != 0 render locally and push their stripes with one 3-D copy-engine transfer (content rect only), rank 0 in place")
                                   + f"; commit (4-byte all-reduce) every {C * SUB} frames")}
        del targets
        if path == "raycast":
            _native.call("rt_raycast_set_view_refit", args.view_refit)
        if store is not None:
            torch.cuda.synchronize()
            barrier_sync(world)
            store.close()
    res["note"] = ("configs[3] as BASELINE.json words it (image-space tiles of one 4K frame over N GPUs + gather), beside the frames partition of the "
                   "primary line.  What does not shrink with N: the per-frame projection + tightening of the BVH (ray cast) resp. vertex shading + "
                   "setup of all triangles (raster) are replicated on every rank.  The orbit is replayed as CUDA graphs: launched eagerly from Python "
                   "the loop is host-bound from N = 4 on (~40 us of Python and driver calls per frame).")
    return res


# ---------------------------------------------------------------------------------------------------------
# configs[4]: 10M-triangle instanced scene, 256-frame orbit, 1080p, raster + ray cast, frames k = rank (mod N)
# ---------------------------------------------------------------------------------------------------------

def bench_config4(args, rank, world, base_rows):
    import torch
    import rendering as ren
    from rendertoy_b200 import scenes
    t0 = time.perf_counter()
    vb = scenes.instanced_device(ren, base_rows, grid=10, scale=0.1, seed=1)
    n_tris = vb.shape[0] // 3
    torch.cuda.synchronize()
    gen_s = time.perf_counter() - t0
    W, H = RAS_W, RAS_H
    per_rank = max(SUB * max(1, args.commit_every), (ORBIT // world) // (SUB * max(1, args.commit_every)) * SUB * max(1, args.commit_every))
    sub = argparse.Namespace(**vars(args))
    sub.steps, sub.warmup = max(2, min(args.steps, 4)), 1
    ras = bench_raster(sub, rank, world, None, vb, W, H, frames=per_rank, full=False, n_tris=n_tris)
    ray = bench_raycast(sub, rank, world, None, vb, W, H, lesson=8, frames=per_rank, full=False, n_tris=n_tris)
    hbm, src = peaks()
    ras_bytes = 3 * n_tris * 32 + W * H * 20
    ray_bytes = 4 * W * H + n_tris * 96 + (2 * n_tris - 1) * 32
    return {"workload": "configs[4]: synthetic 10M-triangle instanced dragon scene (10x10 instances of dragon100k, flattened), 256-frame orbit, "
                        "1920x1080, lesson08 camera, raster + ray cast, frames k = rank (mod N), gathered to rank 0",
            "triangles": n_tris, "n_gpus": world, "frames_per_rank_per_step": per_rank, "steps": sub.steps, "scene_generation_s": gen_s,
            "raster": {"value": ras["value"], "unit": "Mtris/s", "ms_per_frame_per_rank": ras["frame_ms"], "timed_region_ms_per_rank": ras["ms_ranks"],
                       "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": ras_bytes, "achieved": ras_bytes / (ras["frame_ms"] * 1e-3) / 1e9,
                                    "peak": hbm, "unit": "GB/s", "frac": ras_bytes / (ras["frame_ms"] * 1e-3) / 1e9 / hbm, "peak_source": src},
                       "gather": ras["gather"]},
            "raycast": {"value": ray["value"], "unit": "Mrays/s", "ms_per_frame_per_rank": ray["kernel_ms"], "timed_region_ms_per_rank": ray["ms_ranks"],
                        "roofline": {"bound": "hbm", "algorithmic_bytes_per_frame": ray_bytes, "achieved": ray_bytes / (ray["kernel_ms"] * 1e-3) / 1e9,
                                     "peak": hbm, "unit": "GB/s", "frac": ray_bytes / (ray["kernel_ms"] * 1e-3) / 1e9 / hbm, "peak_source": src,
                                     "note": "scene (nodes + leaves, 1.6 GB) > L2: every frame streams what its rays touch; an upper bound on the "
                                             "compulsory bytes, most rays touch a small part of the tree"},
                        "gather": ray["gather"], "traversal": "per-lane 3-D walk (above 2^18 triangles the screen-space projection pass costs more than it saves)"}}


# ---------------------------------------------------------------------------------------------------------
# CPU oracle timing (cpu_baseline and --impl reference)
# ---------------------------------------------------------------------------------------------------------

_CPU_BVH = {}


def cpu_threads():
    """All host cores, explicitly: torchrun hands its workers OMP_NUM_THREADS=1."""
    import oracle
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    return oracle.set_threads(n)


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
    """Times the CPU oracle on this box's host cores: same metric, unit and `config` as our arm, one frame per step."""
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return
    import oracle
    from rendertoy_b200 import scenes
    oracle.build()
    rows = scenes.dragon(N_TRIS)
    cores = cpu_threads()
    if args.path == "raster":
        fn, metric, unit, units, config = cpu_raster, METRIC_RAS, "Mtris/s", N_TRIS, CONFIG_RAS
        sample = "1 frame of 1920x1080 x 100000 triangles per step (the orbit's frame k = step)"
        kind_note = "oracle port of the reference pipeline (rendering/_raster.py kernels restated in C + OpenMP; pyopencl is not installable here)"
    else:
        fn, metric, unit, units, config = cpu_raycast, METRIC_RAY, "Mrays/s", RAY_W * RAY_H, CONFIG_RAY
        sample = "1 frame of 3840x2160 (8.29 Mrays) per step (the orbit's frame k = step)"
        kind_note = "CPU BVH of oracle/raycast_oracle.c: the reference has NO ray caster (rendering/_raycaster.py:35-36 is `pass`)"
    fn(rows, 1, 0)   # builds the CPU BVH (excluded, as on the GPU side) and faults everything in
    for s in range(args.warmup):
        fn(rows, 1, s)
    t0 = time.perf_counter()
    vals = [fn(rows, 1, args.warmup + s)[0] for s in range(args.steps)]
    wall = time.perf_counter() - t0
    value = units * args.steps / wall / 1e6
    print(json.dumps({
        "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
        "run": {"note": kind_note, "frames_per_step": 1, "omp_threads": cores, "per_step": vals},
        "cpu_baseline": {"value": value, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------------------

def assemble_ray(args, world, r):
    hbm_peak, peak_src = peaks()
    facts = ncu_fact("raycast_frame") or ncu_fact("raycast_kernel")
    W, H = RAY_W, RAY_H
    alg_bytes = 4 * W * H + N_TRIS * 96 + (2 * N_TRIS - 1) * 32     # what the timed launch pair moves: BGRA8 out + the scene once (SURVEY.md 8d less the id/t planes the timed call does not write)
    kernel_ms = r["kernel_ms"]
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    import torch
    sm_count = torch.cuda.get_device_properties(0).multi_processor_count
    clock_hz = (r["clocks"]["sm_max_mhz"] or 1965) * 1e6
    fp32_peak = sm_count * 128 * 2 * clock_hz / 1e12
    nodes, tests, rays = r["stats"]
    flops = nodes * 10 + tests * 45
    issue_peak = sm_count * 4 * clock_hz                      # warp instructions per second: one per scheduler and cycle
    winstr = facts.get("warp_instructions_per_frame")
    issue = None
    if winstr:
        issue = {"bound": "issue", "warp_instructions_per_frame": winstr, "achieved": winstr / (kernel_ms * 1e-3) / 1e9, "peak": issue_peak / 1e9,
                 "unit": "G warp-instr/s", "frac": winstr / (kernel_ms * 1e-3) / issue_peak,
                 "issue_slots_busy_pct_ncu": facts.get("issue_slots_busy_pct"), "cycles_per_frame_at_peak": winstr / (sm_count * 4),
                 "source": facts.get("source"),
                 "note": "THE BINDING ROOFLINE: executed warp instructions of project_kernel + raycast_kernel for one frame of this orbit (ncu, "
                         "static for a given scene and camera) over the live event-timed launch duration, against one instruction per "
                         "scheduler and cycle (4 schedulers x SMs x max clock)"}
    hbm = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
           "algorithmic_bytes_per_launch": alg_bytes,
           "algorithmic_bytes_note": "4 B/ray BGRA8 written + T x 96 B leaves + (2T-1) x 32 B nodes read once (SURVEY.md 8d; the id and t planes of its "
                                     "12 B/ray are not written by the timed call and are not counted)",
           "note": "HBM does not bind this kernel: BVH and its per-frame screen-space copy are L2-resident, DRAM sees the frame"}
    common = {"traffic": facts.get("dram_bytes_per_launch"), "kernel": "raycast_kernel<8> (+ project_kernel and view_refit_kernel, same launch sequence)",
              "kernel_ms": kernel_ms, "kernel_ms_alone": r["kernel_ms_alone"],
              "kernel_ms_note": "kernel_ms = timed region / launch sequences in it (frames overlap on the streams named in run); "
                                "kernel_ms_alone = one frame's launch sequence with nothing else on the GPU",
              "fp32": {"achieved_tflops": flops / (kernel_ms * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                       "frac": flops / (kernel_ms * 1e-3) / 1e12 / fp32_peak,
                       "inner_node_visits_per_ray": nodes / (W * H), "triangle_tests_per_ray": tests / (W * H),
                       "rays_traced_fraction": rays / (W * H),
                       "flop_model": "per voting lane: 10 float compares per inner node (two screen rectangles + depth bound each) + 45 flop "
                                     "per Moller-Trumbore test; peak counts FMA as 2 flop, this kernel is compiled -fmad=false for bit-exact parity"}}
    if issue:
        # the stated roofline is the unit that binds: instruction issue.  (The contract's "hbm" figure is kept under `hbm`.)
        ray_roofline = {**{k: issue[k] for k in ("bound", "achieved", "peak", "unit", "frac")}, **common, "issue": issue, "hbm": hbm,
                        "peak_source": "4 warp schedulers x SM count x max SM clock (one warp instruction per scheduler and cycle)",
                        "note": "neither HBM nor the tensor cores bind a BVH walk (no contraction: tensor pipe 0 % in profiles/r02_inventory.txt); "
                                "the binding unit is instruction issue, so that is the roofline stated here; the HBM figure of the contract is under `hbm`"}
    else:
        ray_roofline = {**hbm, **common}
    out = {
        "metric": METRIC_RAY, "value": r["value"], "unit": "Mrays/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": r["ms"] / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": CONFIG_RAY,
        "run": {"frames_per_rank_per_step": r["frames"], "partition": "frames k = rank (mod N); " + r["gather"],
                "l2": f"each rank cycles {SUB} distinct 33 MB frame targets (265 MB > L2); mesh + BVH (~21 MB) stay L2-resident by design, "
                      "as they are reused every frame",
                "streams": f"frames alternate over {args.raycast_streams} CUDA streams", "bvh_build_excluded": True,
                "timed_region_ms_per_rank": r["ms_ranks"], "view_refit_passes": args.view_refit,
                **({"gather_verified": "every rank's locally rendered frames of the last sub-batch == its slots of rank 0's frame store, bit for bit",
                    "gather_bytes_per_step_per_rank": r["push_bytes_step"], "full_frame_bytes_per_step_per_rank": 4 * W * H * r["frames"]}
                   if r["gather_verified"] else {})},
        "roofline": ray_roofline,
        "frame_filling": r["frame_filling"],
        "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"],
    }
    return out


def assemble_ras(args, world, r):
    hbm_peak, peak_src = peaks()
    facts = ncu_fact("raster_frame")
    alg_bytes = 3 * N_TRIS * 32 + RAS_W * RAS_H * 20                               # SURVEY.md section 8(d), per frame
    achieved = alg_bytes / (r["frame_ms"] * 1e-3) / 1e9
    import torch
    sm_count = torch.cuda.get_device_properties(0).multi_processor_count
    issue_peak = sm_count * 4 * 1965e6
    winstr = facts.get("warp_instructions_per_frame")
    issue = None
    if winstr:
        issue = {"bound": "issue", "warp_instructions_per_frame": winstr, "achieved": winstr / (r["frame_ms"] * 1e-3) / 1e9, "peak": issue_peak / 1e9,
                 "unit": "G warp-instr/s", "frac": winstr / (r["frame_ms"] * 1e-3) / issue_peak, "source": facts.get("source"),
                 "note": "executed warp instructions of the four kernels of one cfg2 frame (ncu) over the live per-frame time; the frame's fixed, "
                         "pixel-proportional part (depth clear + resolve of 2.07 M pixels + the mostly idle coverage grid) is ~31 us of the ~48 "
                         "whatever the triangle count (measured with 2k .. 200k triangles, DESIGN.md 4.3)"}
    return {
        "metric": METRIC_RAS, "value": r["value"], "unit": "Mtris/s", "ms_per_step": r["ms"] / args.steps, "scaling": "weak", "config": CONFIG_RAS,
        "run": {"frames_per_rank_per_step": r["frames"], "partition": "frames k = rank (mod N); " + r["gather"],
                "l2": f"{SUB} independent raster targets per rank (~300 MB of key/colour/record buffers > L2)",
                "streams": f"one CUDA stream per frame target ({r['streams']})", "timed_region_ms_per_rank": r["ms_ranks"],
                **({"gather_verified": "every rank's locally rendered frames of the last sub-batch == its slots of rank 0's frame store, bit for bit",
                    "gather_bytes_per_step_per_rank": r["push_bytes_step"], "full_frame_bytes_per_step_per_rank": 4 * RAS_W * RAS_H * r["frames"]}
                   if r["gather_verified"] else {})},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                     "traffic": facts.get("dram_bytes_per_launch"),
                     "peak_source": peak_src, "unit_of_work": "one frame = 4 kernels (depth clear, raster_kernel, "
                     "coverage_kernel, resolve_kernel; the colour clear is folded into the resolve); algorithmic bytes are defined per frame",
                     "frame_ms": r["frame_ms"], "frame_ms_alone": r["frame_ms_alone"], "algorithmic_bytes_per_frame": alg_bytes, "issue": issue},
        "e2e": r["e2e"], "gpu_launches": r["launches"],
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--path", default="raycast", choices=["raycast", "raster"], help="which half of the metric is the primary line")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-tiles", action="store_true", help="skip the image-space tile partition (configs[3] as written)")
    ap.add_argument("--no-config4", action="store_true", help="skip configs[4] (10M triangles)")
    ap.add_argument("--only", default=None, choices=["raycast", "raster", "tiles", "config4"], help="dev: run one section only")
    ap.add_argument("--raycast-streams", type=int, default=4, help="raycast frames alternate over this many CUDA streams")
    ap.add_argument("--raster-streams", type=int, default=1, help="1: one CUDA stream per raster frame target (default), 0: single stream")
    ap.add_argument("--commit-every", type=int, default=4, help="N>1: sub-batches of 8 frames per rank between two commits (4-byte all-reduce)")
    ap.add_argument("--dense-gather", dest="sparse", action="store_false",
                    help="--gather copy: push whole frames instead of the rect that can differ from the clear colour")
    ap.add_argument("--readback", default="rect", choices=["rect", "tiles"],
                    help="e2e: rect = copy-engine D2H of the content rect (rt_copy_rect); tiles = rt_push_tiles, a kernel storing the non-clear "
                         "32x32 tiles into the pinned host frame")
    ap.add_argument("--dense-readback", dest="sparse_readback", action="store_false",
                    help="e2e: read whole frames back instead of the rect that can differ from the clear colour")
    ap.add_argument("--view-refit", type=int, default=None, help="tightening passes over the screen-space nodes (rt_raycast_set_view_refit)")
    ap.add_argument("--raster-push", default="tiles", choices=["tiles", "rect"],
                    help="--raster-gather copy: tiles = rt_push_tiles (kernel on the producer, non-clear 32x32 tiles); rect = rt_copy_rect "
                         "(copy engine, Raster.content_rect)")
    ap.add_argument("--raster-gather", default="copy", choices=["peer", "copy", "nccl"],
                    help="N>1, raster frames: copy = render locally, push Raster.content_rect with the copy engine; peer = the kernels "
                         "store into rank 0's frame store")
    ap.add_argument("--gather", default="copy", choices=["peer", "copy", "nccl"],
                    help="N>1, raycast frames: copy = ranks render locally and a copy engine pushes each finished frame into rank 0's "
                         "IPC-mapped frame store while the next frames trace; peer = the kernels store straight into that frame store "
                         "over NVLink (fused); nccl = send/recv gather")
    ap.add_argument("--tiles-graphs", type=int, default=1, help="tile partition: 1 = replay the orbit as CUDA graphs (default), 0 = eager launches from Python")
    ap.add_argument("--tiles-view-refit", type=int, default=None, help="tile partition: view-node tightening iterations (default: as --view-refit)")
    ap.add_argument("--tiles-gather", default="auto", choices=["auto", "peer", "copy"],
                    help="N>1, tile partition: how stripes reach rank 0's frame; auto = copy for ray-cast frames, peer for raster frames "
                         "(measured at N=2 and 8: 123/174 vs 108/149 Grays/s, 2120/2731 vs 1392/1349 Mtris/s)")
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
    rank, world, local = dist_setup()
    from rendertoy_b200 import _native, scenes
    import rendering as ren
    if args.view_refit is None:
        args.view_refit = DEFAULT_VIEW_REFIT
    _native.call("rt_raycast_set_view_refit", args.view_refit)
    rows = scenes.dragon(N_TRIS)
    vb = upload(ren, rows)
    want = (lambda name: args.only in (None, name))
    ray = ras = tiles = cfg4 = None
    if want("raycast"):
        ray = assemble_ray(args, world, bench_raycast(args, rank, world, rows, vb))
    if want("raster"):
        ras = assemble_ras(args, world, bench_raster(args, rank, world, rows, vb))
    if want("tiles") and not args.no_tiles:
        tiles = bench_tiles(args, rank, world, rows, vb)
    if want("config4") and not args.no_config4:
        cfg4 = bench_config4(args, rank, world, rows)
    if rank == 0:
        if not args.no_cpu_baseline and args.only is None and world == 1:     # the CPU legs belong to the N = 1 line
            import oracle
            oracle.build()
            cores = cpu_threads()
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
        if args.only is not None:
            primary = {"only": args.only, "n_gpus": world, "raycast": ray, "raster": ras, "tiles": tiles, "config4": cfg4}
        else:
            primary, secondary = (ray, ras) if args.path == "raycast" else (ras, ray)
            if args.path == "raster":
                for k in ("n_gpus", "steps", "warmup", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "clocks"):
                    primary.setdefault(k, ray.get(k))
            primary["secondary"] = secondary
            if tiles is not None:
                primary["tiles"] = tiles
            if cfg4 is not None:
                primary["config4"] = cfg4
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(primary) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


DEFAULT_VIEW_REFIT = 4   # iterations inside the one view_refit_kernel launch; measured on B200 (profiles/r02f_refit_iterations.txt): 0 -> 70.9, 2 -> 75.1, 4 -> 76.2, 8 -> 74.4, 16 -> 71.8 Grays/s on the cfg4 orbit


if __name__ == "__main__":
    main()
