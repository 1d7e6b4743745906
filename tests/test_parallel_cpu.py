"""Partition logic and the one collective, on CPU: world_size-2 gloo processes."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rendertoy_b200 import parallel


def test_tile_rects_cover_frame_exactly():
    for (w, h, world) in [(3840, 2160, 8), (1920, 1080, 4), (333, 211, 3), (64, 10, 2)]:
        seen = np.zeros((h, w), np.int32)
        for r in range(world):
            for (x0, y0, tw, th) in parallel.tile_rects(w, h, r, world):
                assert y0 % 8 == 0 and x0 == 0
                seen[y0:y0 + th, x0:x0 + tw] += 1
        assert (seen == 1).all()


def test_frame_indices_partition():
    for n, world in [(256, 8), (10, 4), (3, 8)]:
        allk = sorted(k for r in range(world) for k in parallel.frame_indices(n, r, world))
        assert allk == list(range(n))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, w, h, n_frames, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # tiles: every rank fills its own bands with (rank+1), rank 0 must end up with the full pattern
        frame = torch.zeros((h, w), dtype=torch.int32)
        for (x0, y0, tw, th) in parallel.tile_rects(w, h, rank, world):
            frame[y0:y0 + th, x0:x0 + tw] = rank + 1
        parallel.gather_tiles(frame, w, h)
        # frames: frame k is filled with k
        mine = parallel.frame_indices(n_frames, rank, world)
        local = torch.stack([torch.full((h, w), k, dtype=torch.int32) for k in mine]) if mine else torch.zeros((0, h, w), dtype=torch.int32)
        allf = torch.full((n_frames, h, w), -1, dtype=torch.int32) if rank == 0 else None
        parallel.gather_frames(local, allf, n_frames)
        if rank == 0:
            expect = torch.zeros((h, w), dtype=torch.int32)
            for r in range(world):
                for (x0, y0, tw, th) in parallel.tile_rects(w, h, r, world):
                    expect[y0:y0 + th, x0:x0 + tw] = r + 1
            ok = bool((frame == expect).all()) and all(bool((allf[k] == k).all()) for k in range(n_frames))
            out.put(ok)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_gather_world_size_2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 96, 200, 5, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(100)
        assert p.exitcode == 0
    assert out.get(timeout=5) is True


# ---- sparse frame movement (cover_rect / SparseFrameCopier) -------------------------------------------------------

def _fake_copy_rect(name, dst, dst_pitch, src, src_pitch, width_bytes, rows, stream):
    """What rt_copy_rect does (cudaMemcpy2DAsync), on host memory."""
    import ctypes
    assert name == "rt_copy_rect" and dst_pitch >= width_bytes and src_pitch >= width_bytes
    for r in range(rows):
        ctypes.memmove(dst + r * dst_pitch, src + r * src_pitch, width_bytes)


def test_cover_rect_cases():
    W, H = 200, 100
    assert parallel.cover_rect(None, (5, 1, 4, 9), W, H) is None                 # nothing before, nothing now
    assert parallel.cover_rect(None, (40, 10, 70, 20), W, H) == (32, 10, 95, 20)  # columns widen to 32-px lines
    assert parallel.cover_rect((40, 10, 70, 20), (5, 1, 4, 9), W, H) == (32, 10, 95, 20)   # the old content must be cleared
    assert parallel.cover_rect((0, 0, 10, 10), (150, 50, 199, 99), W, H) == (0, 0, 199, 99)
    assert parallel.cover_rect(None, (190, 0, 500, 500), W, H) == (160, 0, 199, 99)  # clipped; last line partial
    assert parallel.cover_rect(None, (-20, -5, 3, 3), W, H) == (0, 0, 31, 3)


def test_sparse_copier_keeps_destinations_exact(monkeypatch):
    """Random frames that are clear outside a random content rect, cycled through two persistent destinations:
    after every copy the destination must equal the source frame everywhere, with far fewer bytes moved."""
    from rendertoy_b200 import _native
    monkeypatch.setattr(_native, "call", _fake_copy_rect)
    rng = np.random.default_rng(7)
    W, H = 333, 97
    copier = parallel.SparseFrameCopier(W, H)
    dests = [np.zeros((H, W), np.uint32) for _ in range(2)]
    moved = 0
    for it in range(200):
        src = np.zeros((H, W), np.uint32)
        kind = it % 10
        if kind == 0:
            content = (5, 5, 4, 4)                                               # empty: nothing can be hit
        elif kind == 1:
            content = (0, 0, W - 1, H - 1)                                       # no bound known: the whole frame
        else:
            x0, x1 = sorted(rng.integers(0, W, 2)); y0, y1 = sorted(rng.integers(0, H, 2))
            content = (int(x0), int(y0), int(x1), int(y1))
        if content[2] >= content[0]:
            x0, y0, x1, y1 = content
            src[y0:y1 + 1, x0:x1 + 1] = rng.integers(0, 2 ** 32, (y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint32)
        d = dests[it % 2]
        moved += copier.copy(it % 2, d.ctypes.data, src.ctypes.data, content, None)
        assert np.array_equal(d, src), f"iteration {it}: destination differs from the frame"
    assert moved == copier.bytes_moved and moved < 0.8 * 200 * W * H * 4


def test_screen_bounds_native_matches_numpy(ren):
    """rt_raycast_screen_bounds (host-only C) against the numpy formulation it replaced, lesson cameras + degenerate ones."""
    import ctypes
    from rendering._raycaster import Raycaster, camera_frame
    from rendertoy_b200 import scenes
    rc = Raycaster.__new__(Raycaster)
    rc.scene_lo, rc.scene_hi = np.array([-0.5, -0.31, -0.22]), np.array([0.5, 0.27, 0.24])
    rc.scene_extent = 1.0
    rc._lo3, rc._hi3 = (ctypes.c_double * 3)(*rc.scene_lo), (ctypes.c_double * 3)(*rc.scene_hi)
    rc._chunk_lo, rc._chunk_hi, rc._n_chunks = rc._lo3, rc._hi3, 1         # one chunk: the scene box itself

    def numpy_bounds(camera, W, H, lo=None, hi=None):
        cam = np.asarray(camera, np.float64).reshape(4, 3)
        o, basis = cam[0], cam[1:4].T
        lo, hi = (rc.scene_lo, rc.scene_hi) if lo is None else (lo, hi)
        corners = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])
        try:
            abc = np.linalg.solve(basis, (corners - o).T)
        except np.linalg.LinAlgError:
            return None
        if not np.all(np.isfinite(abc)) or abc[2].min() <= 1e-6:
            return None
        px = (abc[0] / abc[2] + 1.0) * (W * 0.5) - 0.5
        py = (1.0 - abc[1] / abc[2]) * (H * 0.5) - 0.5
        return (int(max(0, np.floor(px.min()) - 2)), int(max(0, np.floor(py.min()) - 2)),
                int(min(W - 1, np.ceil(px.max()) + 2)), int(min(H - 1, np.ceil(py.max()) + 2)))

    for k in range(64):
        for lesson, (W, H) in ((6, (3840, 2160)), (8, (1920, 1080)), (6, (333, 211))):
            world, view, proj = scenes.lesson_camera(ren, lesson, 0.37 * k, W, H)
            cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
            assert rc.screen_bounds(cam, W, H) == numpy_bounds(cam, W, H)
    # several chunk boxes (Raycaster: bounds of 64 chunks of the sorted leaves): the union of the single rectangles
    rng = np.random.default_rng(4)
    centres = rng.uniform(-0.4, 0.4, (7, 3))
    clo, chi = centres - rng.uniform(0.01, 0.1, (7, 3)), centres + rng.uniform(0.01, 0.1, (7, 3))
    rc._chunk_lo, rc._chunk_hi, rc._n_chunks = (ctypes.c_double * 21)(*clo.ravel()), (ctypes.c_double * 21)(*chi.ravel()), 7
    for k in range(16):
        W, H = 1920, 1080
        world, view, proj = scenes.lesson_camera(ren, 8 if k % 2 else 6, 0.41 * k, W, H)
        cam = camera_frame(np.array(view, dtype=ren.float4x4), np.array(proj, dtype=ren.float4x4), np.array(world, dtype=ren.float4x4))
        singles = [numpy_bounds(cam, W, H, clo[i], chi[i]) for i in range(7)]
        assert all(s_ is not None for s_ in singles)
        vis = [s_ for s_ in singles if s_[2] >= s_[0] and s_[3] >= s_[1]]
        want = (min(s_[0] for s_ in vis), min(s_[1] for s_ in vis), max(s_[2] for s_ in vis), max(s_[3] for s_ in vis))
        assert rc.screen_bounds(cam, W, H) == want
    rc._chunk_lo, rc._chunk_hi, rc._n_chunks = rc._lo3, rc._hi3, 1
    inside = np.array([0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1], np.float32)           # eye inside the box
    assert rc.screen_bounds(inside, 640, 480) is None
    singular = np.array([0, 0, 3, 1, 0, 0, 1, 0, 0, 0, 0, -1], np.float32)        # U == V
    assert rc.screen_bounds(singular, 640, 480) is None
    nan = np.array([np.nan, 0, 3, 1, 0, 0, 0, 1, 0, 0, 0, -1], np.float32)
    assert rc.screen_bounds(nan, 640, 480) is None
    far = np.array([1e6, 0, 3, 1, 0, 0, 0, 1, 0, 0, 0, -1], np.float32)           # scene far off to the side: clamped, maybe empty
    r = rc.screen_bounds(far, 640, 480)
    assert r is None or (0 <= r[0] and r[2] <= 639 and 0 <= r[1] and r[3] <= 479)


def _sparse_worker(rank, world, port, shm_name, w, h, F, steps, out):
    """The bench's sparse-push protocol with the CUDA pieces swapped for host ones: rank 0's double-buffered frame store is a
    shared-memory block every rank maps (CUDA IPC in the product), rt_copy_rect is the host copy above, commit is a gloo
    all-reduce.  Frames: k = rank (mod world); slot (s % 2) * n_frames + k; only cover_rect travels."""
    from multiprocessing import shared_memory
    from rendertoy_b200 import _native
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        _native.call = _fake_copy_rect
        n_frames = F * world
        store = np.ndarray((2 * n_frames, h, w), np.uint32, buffer=shm.buf)
        base = store.ctypes.data
        copier = parallel.SparseFrameCopier(w, h)
        mine = parallel.frame_indices(n_frames, rank, world)
        flag = torch.zeros(1, dtype=torch.int32)
        ok = True

        def frame(s, k):                                     # what "rendering" frame k of step s produces, on any rank
            rng = np.random.default_rng(1000 * s + k)
            f = np.zeros((h, w), np.uint32)
            x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2))
            if (s + k) % 7 == 0:
                return f, (3, 3, 2, 2)                       # nothing visible this frame
            f[y0:y1 + 1, x0:x1 + 1] = rng.integers(1, 2 ** 32, (y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint32)
            return f, (int(x0), int(y0), int(x1), int(y1))

        for s in range(steps):
            for k in mine:
                f, content = frame(s, k)
                slot = (s % 2) * n_frames + k
                if rank == 0:
                    store[slot] = f                          # rank 0 renders in place (whole frame, clear pixels included)
                else:
                    copier.copy(slot, base + slot * w * h * 4, f.ctypes.data, content, None)
            dist.all_reduce(flag)                            # commit: every rank's pushes of this step are done
            if rank == 0:
                for k in range(n_frames):
                    ok &= bool(np.array_equal(store[(s % 2) * n_frames + k], frame(s, k)[0]))
            dist.barrier()                                   # (the product re-uses a slot only two commits later)
        if rank == 0:
            out.put(ok)
        else:
            out.put(copier.bytes_moved < 0.8 * steps * F * w * h * 4)
    finally:
        shm.close()
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_sparse_push_protocol_world_size_2_gloo():
    from multiprocessing import shared_memory
    w, h, F, steps = 96, 40, 3, 6
    shm = shared_memory.SharedMemory(create=True, size=2 * F * 2 * w * h * 4)
    try:
        np.ndarray((2 * F * 2, h, w), np.uint32, buffer=shm.buf)[:] = 0      # the store starts cleared
        ctx = mp.get_context("spawn")
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_sparse_worker, args=(r, 2, port, shm.name, w, h, F, steps, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(100)
            assert p.exitcode == 0
        assert out.get(timeout=5) is True and out.get(timeout=5) is True
    finally:
        shm.close()
        shm.unlink()


# ---- image-space partition of one frame: stripes -------------------------------------------------------------------

def _fake_native(name, *a):
    """Host stand-ins for the two copy entry points (semantics as documented in include/rendertoy_b200.h)."""
    import ctypes
    if name == "rt_copy_rect":
        return _fake_copy_rect(name, *a)
    assert name == "rt_copy_stripes"
    dst, src, pitch, x_bytes, width_bytes, y0, y1, rows, mod, rem, stream = a
    assert x_bytes + width_bytes <= pitch and rows >= 1 and 0 <= rem < mod
    for y in range(y0, y1 + 1):
        if (y // rows) % mod == rem:
            ctypes.memmove(dst + y * pitch + x_bytes, src + y * pitch + x_bytes, width_bytes)


def test_stripes_are_the_tile_rects():
    """stripes_of(rank, world) names exactly the rows tile_rects() hands to that rank."""
    for (w, h, world) in [(3840, 2160, 8), (1920, 1080, 3), (333, 211, 2)]:
        for r in range(world):
            rows, mod, rem = parallel.stripes_of(r, world)
            own = np.zeros(h, bool)
            for (x0, y0, tw, th) in parallel.tile_rects(w, h, r, world):
                own[y0:y0 + th] = True
            assert np.array_equal(own, (np.arange(h) // rows) % mod == rem)


def test_stripe_copier_assembles_the_frame(monkeypatch):
    """Random frames that are clear outside a random content rect, every "rank" pushing only its stripes of the cover rect into
    two persistent destinations: after all ranks pushed, the destination equals the frame; a rank never writes foreign rows."""
    from rendertoy_b200 import _native
    monkeypatch.setattr(_native, "call", _fake_native)
    rng = np.random.default_rng(11)
    W, H, world, band = 333, 211, 3, 16
    copiers = [parallel.SparseFrameCopier(W, H) for _ in range(world)]      # one history per rank, as in the product
    dests = [np.zeros((H, W), np.uint32) for _ in range(2)]
    yy = np.arange(H)
    for it in range(120):
        src = np.zeros((H, W), np.uint32)
        if it % 9 == 0:
            content = (7, 7, 6, 6)
        elif it % 9 == 1:
            content = (0, 0, W - 1, H - 1)
        else:
            x0, x1 = sorted(rng.integers(0, W, 2)); y0, y1 = sorted(rng.integers(0, H, 2))
            content = (int(x0), int(y0), int(x1), int(y1))
        if content[2] >= content[0]:
            x0, y0, x1, y1 = content
            src[y0:y1 + 1, x0:x1 + 1] = rng.integers(1, 2 ** 32, (y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint32)
        d = dests[it % 2]
        for r in range(world):
            before = d.copy()
            copiers[r].copy_stripes(it % 2, d.ctypes.data, src.ctypes.data, content, (band, world, r), None)
            foreign = (yy // band) % world != r
            assert np.array_equal(d[foreign], before[foreign]), f"iteration {it}: rank {r} wrote foreign rows"
        assert np.array_equal(d, src), f"iteration {it}: destination differs from the frame"
    assert sum(c.bytes_moved for c in copiers) < 0.8 * 120 * W * H * 4


def _stripe_worker(rank, world, port, shm_name, w, h, frames, out):
    """The tile partition's protocol with host stand-ins: rank 0's frame ring is a shared-memory block (CUDA IPC in the product),
    every rank "renders" the whole synthetic frame locally and pushes only its stripes of the cover rect (rank 0 writes its
    stripes in place), a gloo all-reduce commits; rank 0 then holds the complete frame."""
    from multiprocessing import shared_memory
    from rendertoy_b200 import _native
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shm = shared_memory.SharedMemory(name=shm_name)
    try:
        _native.call = _fake_native
        ring = 4
        store = np.ndarray((ring, h, w), np.uint32, buffer=shm.buf)
        copier = parallel.SparseFrameCopier(w, h)
        stripes = parallel.stripes_of(rank, world, band=8)
        mine = (np.arange(h) // 8) % world == rank
        flag = torch.zeros(1, dtype=torch.int32)
        ok = True

        def frame(f):
            rng = np.random.default_rng(77 + f)
            img = np.zeros((h, w), np.uint32)
            if f % 5 == 0:
                return img, (2, 2, 1, 1)
            x0, x1 = sorted(rng.integers(0, w, 2)); y0, y1 = sorted(rng.integers(0, h, 2))
            img[y0:y1 + 1, x0:x1 + 1] = rng.integers(1, 2 ** 32, (y1 - y0 + 1, x1 - x0 + 1), dtype=np.uint32)
            return img, (int(x0), int(y0), int(x1), int(y1))

        for f in range(frames):
            img, content = frame(f)
            slot = f % ring
            if rank == 0:
                store[slot][mine] = img[mine]                 # in place: its own stripes, clear pixels included
            else:
                copier.copy_stripes(slot, store[slot].ctypes.data, img.ctypes.data, content, stripes, None)
            dist.all_reduce(flag)                              # commit
            if rank == 0:
                ok &= bool(np.array_equal(store[slot], img))
            dist.barrier()
        out.put(ok if rank == 0 else copier.bytes_moved < 0.5 * frames * w * h * 4)
    finally:
        shm.close()
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_stripe_partition_protocol_world_size_2_gloo():
    from multiprocessing import shared_memory
    w, h, frames = 96, 44, 14
    shm = shared_memory.SharedMemory(create=True, size=4 * w * h * 4)
    try:
        np.ndarray((4, h, w), np.uint32, buffer=shm.buf)[:] = 0
        ctx = mp.get_context("spawn")
        out = ctx.Queue()
        port = _free_port()
        procs = [ctx.Process(target=_stripe_worker, args=(r, 2, port, shm.name, w, h, frames, out)) for r in range(2)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(100)
            assert p.exitcode == 0
        assert out.get(timeout=5) is True and out.get(timeout=5) is True
    finally:
        shm.close()
        shm.unlink()
