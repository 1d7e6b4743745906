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
