"""Multi-GPU partitioning for the two hot paths (the reference is single-device,
rendering/_core.py:10-11).

One process per GPU (torchrun), mesh + BVH replicated on every rank, no data-path exchange except ONE
collective: finished BGRA8 framebuffer pieces are gathered to rank 0 (NCCL over NVLink; gloo in CPU tests).

  tile_rects(W, H, rank, world)        image-space partition of one large frame: 8-aligned row bands, round-robin
  frame_indices(n_frames, rank, world) animation batches: rank r renders frames k = r (mod world)
  gather_tiles / gather_frames         the collective (grouped isend/irecv, i.e. ncclSend/ncclRecv: NCCL has no
                                       native gather)
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

BAND = 64  # rows per band: multiple of the traversal's 8x4 warp tile, small enough to balance a frame


def tile_rects(width: int, height: int, rank: int, world: int, band: int = BAND) -> List[Tuple[int, int, int, int]]:
    """(x0, y0, w, h) row bands owned by `rank`: band b belongs to rank b % world.  Bands interleave so every
    rank sees a similar mix of background and mesh."""
    rects = []
    for b, y0 in enumerate(range(0, height, band)):
        if b % world == rank:
            rects.append((0, y0, width, min(band, height - y0)))
    return rects


def frame_indices(n_frames: int, rank: int, world: int) -> List[int]:
    return list(range(rank, n_frames, world))


def _is_dist():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def gather_tiles(frame: torch.Tensor, width: int, height: int, band: int = BAND, dst: int = 0):
    """frame: (H, W) int32/uint8x4 tensor holding this rank's bands at their final positions.  After the call
    rank `dst` holds every band.  Each band is one send/recv, all grouped in a single batch."""
    if not _is_dist():
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    for b, y0 in enumerate(range(0, height, band)):
        owner = b % world
        rows = frame[y0:min(y0 + band, height)]
        if owner == dst:
            continue
        if rank == owner:
            ops.append(dist.P2POp(dist.isend, rows, dst))
        elif rank == dst:
            ops.append(dist.P2POp(dist.irecv, rows, owner))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


def gather_frames(local_frames: torch.Tensor, all_frames: torch.Tensor, n_frames: int, dst: int = 0):
    """Animation batch: local_frames (n_local, H, W) holds frames rank, rank+world, ...; all_frames
    (n_frames, H, W) on rank `dst` receives every frame at its index."""
    if not _is_dist():
        if all_frames is not None and all_frames.data_ptr() != local_frames.data_ptr():
            all_frames[:local_frames.shape[0]].copy_(local_frames)
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    ops = []
    if rank == dst:
        for j, k in enumerate(frame_indices(n_frames, dst, world)):
            all_frames[k].copy_(local_frames[j])
        for src in range(world):
            if src == dst:
                continue
            for k in frame_indices(n_frames, src, world):
                ops.append(dist.P2POp(dist.irecv, all_frames[k], src))
    else:
        for j, _ in enumerate(frame_indices(n_frames, rank, world)):
            ops.append(dist.P2POp(dist.isend, local_frames[j], dst))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()


class _RawDeviceMemory:
    """Adapter exposing a raw device address to torch through __cuda_array_interface__."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


class FrameStore:
    """n_frames BGRA8 frames of W x H in rank `dst`'s HBM, writable from every rank of the node.

    The fused alternative to gather_frames(): rank `dst` cudaMalloc's the store and exports it with CUDA IPC, the
    other ranks map it, and `frame(k)` returns a torch uint8 view a render target can be built on
    (`rendering.Image(W, H, RGBA, memory=store.frame(k))`).  The ray-cast / resolve kernels then write finished
    pixels straight into rank dst's memory over NVLink while they run; `commit()` (a barrier) is all that remains of
    the collective.  Falls back (ok=False) when IPC mapping is unavailable; callers then use gather_frames()."""

    def __init__(self, n_frames, width, height, dst=0):
        import ctypes
        from . import _native
        self.n_frames, self.width, self.height, self.dst = n_frames, width, height, dst
        self.frame_bytes = width * height * 4
        self.rank = dist.get_rank() if _is_dist() else 0
        self.world = dist.get_world_size() if _is_dist() else 1
        self._native, self._base, self._owner = _native, None, self.rank == dst
        nbytes = self.frame_bytes * n_frames
        handle = [None]
        ok = True
        try:
            if self._owner:
                p = ctypes.c_void_p()
                _native.call("rt_peer_alloc", nbytes, ctypes.byref(p))
                self._base = p.value
                buf = ctypes.create_string_buffer(64)
                _native.call("rt_peer_export", self._base, buf)
                handle = [bytes(buf.raw)]
        except Exception:
            ok = False
        if self.world > 1:
            dist.broadcast_object_list(handle, src=dst)
            if not self._owner and handle[0] is not None:
                try:
                    p = ctypes.c_void_p()
                    _native.call("rt_peer_open", ctypes.create_string_buffer(handle[0], 64), ctypes.byref(p))
                    self._base = p.value
                except Exception:
                    ok = False
            flags = [None] * self.world
            dist.all_gather_object(flags, ok and self._base is not None)
            ok = all(flags)
        self.ok = ok and self._base is not None
        self.memory = torch.as_tensor(_RawDeviceMemory(self._base, nbytes), device="cuda") if self.ok else None
        self._flag = torch.zeros(1, dtype=torch.int32, device="cuda") if self.world > 1 else None

    def frame(self, k):
        """uint8 view (frame_bytes,) of frame k."""
        return self.memory[k * self.frame_bytes:(k + 1) * self.frame_bytes]

    def frames(self):
        """(n_frames, H, W) int32 view -- meaningful on rank `dst` after commit()."""
        return self.memory.view(torch.int32).view(self.n_frames, self.height, self.width)

    def commit(self):
        """Stream-ordered barrier: a 4-byte all-reduce enqueued behind this rank's rendering kernels.  Work enqueued
        after it (on any rank) runs only once every rank's kernels that wrote into the store have finished; the host
        is not blocked, so consecutive batches keep pipelining.  Re-use a slot only two commits later (double-buffer
        the store), so rank `dst` has consumed it in between."""
        if self.world > 1:
            dist.all_reduce(self._flag)

    def close(self):
        if self._base is None:
            return
        self.memory = None
        if self._owner:
            self._native.call("rt_peer_free", self._base)
        else:
            self._native.call("rt_peer_close", self._base)
        self._base = None
