"""Multi-GPU partitioning for the two hot paths (the reference is single-device,
rendering/_core.py:10-11).

One process per GPU (torchrun), mesh + BVH replicated on every rank, no data-path exchange except ONE
collective: finished BGRA8 framebuffer pieces are gathered to rank 0 (NCCL over NVLink; gloo in CPU tests).

  tile_rects(W, H, rank, world)        image-space partition of one large frame: 8-aligned row bands, round-robin
  frame_indices(n_frames, rank, world) animation batches: rank r renders frames k = r (mod world)
  gather_tiles / gather_frames         the collective (grouped isend/irecv, i.e. ncclSend/ncclRecv: NCCL has no
                                       native gather)
  FrameStore                           the same collective over CUDA IPC peer memory: rank 0's store is every rank's
                                       render target (fused stores) or the destination of copy-engine pushes
  cover_rect / SparseFrameCopier       sparse frame movement: a ray-cast frame is the clear colour outside the scene's
                                       projected bounds, so only that pixel rectangle travels (NVLink push, PCIe read-back)
"""
from typing import List, Tuple

import torch
import torch.distributed as dist

BAND = 64  # rows per band: multiple of the traversal's 8x4 warp tile, small enough to balance a frame


def stripes_of(rank: int, world: int, band: int = BAND) -> Tuple[int, int, int]:
    """(rows, mod, rem): the same partition as tile_rects() in the form Raster.set_scissor(stripes=) and
    Raycaster.render(stripes=) take -- one launch per rank and frame instead of one per band."""
    return (band, world, rank)


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


def _empty(r):
    return r is None or r[2] < r[0] or r[3] < r[1]


def cover_rect(prev, cur, width, height, align_px=32):
    """The inclusive pixel rect (x0, y0, x1, y1) that has to be copied so that a destination frame which currently
    holds a frame with content rect `prev` (None: still cleared) becomes the frame with content rect `cur`, given
    that the SOURCE frame is the clear colour everywhere outside `cur`: the union of both (the part of `prev` outside
    `cur` is overwritten with the source's clear pixels), clipped to the frame, columns widened to multiples of
    `align_px` pixels (32 px = one 128-byte line) so the copy engine moves whole lines.  None: nothing to move."""
    if _empty(prev) and _empty(cur):
        return None
    if _empty(prev):
        x0, y0, x1, y1 = cur
    elif _empty(cur):
        x0, y0, x1, y1 = prev
    else:
        x0, y0, x1, y1 = min(prev[0], cur[0]), min(prev[1], cur[1]), max(prev[2], cur[2]), max(prev[3], cur[3])
    x0, y0, x1, y1 = max(0, x0), max(0, y0), min(width - 1, x1), min(height - 1, y1)
    if x1 < x0 or y1 < y0:
        return None
    x0 = x0 // align_px * align_px
    x1 = min(width, (x1 // align_px + 1) * align_px) - 1
    return x0, y0, x1, y1


class SparseFrameCopier:
    """Moves BGRA8 frames that are clear outside a known content rect (what Raycaster.render returns) into destinations
    that persist between frames -- a slot of the frame store, a pinned host frame -- copying only cover_rect(previous
    content of that destination, this frame's content) with one pitched copy-engine transfer (rt_copy_rect).  The
    destination must start cleared (FrameStore does; torch.zeros(...).pin_memory() for host frames) and every frame
    that goes into it must go through the same copier."""

    def __init__(self, width, height):
        from . import _native
        self._native, self.width, self.height = _native, width, height
        self._content = {}          # destination key -> content rect of the frame it holds
        self.bytes_moved = 0

    def copy(self, key, dst_ptr, src_ptr, content, stream):
        """Enqueue on `stream` (raw cudaStream_t).  key: anything hashable naming the destination frame."""
        r = cover_rect(self._content.get(key), content, self.width, self.height)
        self._content[key] = content
        if r is None:
            return 0
        x0, y0, x1, y1 = r
        off, pitch = 4 * (y0 * self.width + x0), 4 * self.width
        self._native.call("rt_copy_rect", dst_ptr + off, pitch, src_ptr + off, pitch, 4 * (x1 - x0 + 1), y1 - y0 + 1, stream)
        n = 4 * (x1 - x0 + 1) * (y1 - y0 + 1)
        self.bytes_moved += n
        return n

    def copy_stripes(self, key, dst_ptr, src_ptr, content, stripes, stream):
        """The same for ONE rank's share of a frame that several ranks write (image-space partition): of the cover rect only the
        rows of the stripes (rows, mod, rem) -- stripe s = y // rows is this rank's iff s % mod == rem -- travel, as one 3-D
        copy-engine transfer (rt_copy_stripes).  dst_ptr / src_ptr: the two full frames (same layout).  Every rank keeps the
        destination's content history for its own stripes, keyed by (key, stripes)."""
        key = ("stripes", key, tuple(stripes))
        r = cover_rect(self._content.get(key), content, self.width, self.height)
        self._content[key] = content
        if r is None:
            return 0
        x0, y0, x1, y1 = r
        rows, mod, rem = stripes
        self._native.call("rt_copy_stripes", dst_ptr, src_ptr, 4 * self.width, 4 * x0, 4 * (x1 - x0 + 1), y0, y1, rows, mod, rem, stream)
        n = 4 * (x1 - x0 + 1) * sum(min(y1, (s + 1) * rows - 1) - max(y0, s * rows) + 1
                                    for s in range(y0 // rows, y1 // rows + 1) if s % mod == rem)
        self.bytes_moved += n
        return n


class TileFrameCopier:
    """Moves BGRA8 frames into destinations that persist between frames by 32x32-pixel tiles (rt_push_tiles): a small kernel on
    the GPU that holds the frame stores only the tiles that hold something other than the clear colour -- or did the last time
    the same destination was written through this copier -- so that afterwards the destination equals the frame.  The
    destination may be another GPU's memory (a FrameStore slot: the multi-GPU gather of frame-filling raster frames) or PINNED
    HOST memory (the read-back of a frame: with unified addressing the kernel's stores cross PCIe directly, and only the
    covered tiles do).  Destinations must start filled with the clear colour (else pass dirty=True for their first frame)."""

    def __init__(self, width, height):
        from . import _native
        self._native, self.width, self.height = _native, width, height
        self._n = int(_native.lib().rt_push_tiles_state_bytes(width, height))
        self._state = {}            # destination key -> (uint8 tensor, one byte per tile; clear colour it refers to)
        self.bytes = torch.zeros(1, dtype=torch.int64, device="cuda")      # bytes stored so far (device counter)
        torch.cuda.current_stream().synchronize()

    def copy(self, key, dst_ptr, src_ptr, stream, clear_px=0, dirty=False):
        st = self._state.get(key)
        if st is None or st[1] != clear_px:
            t = torch.full((self._n,), 1 if (dirty or st is not None) else 0, dtype=torch.uint8, device="cuda")
            torch.cuda.current_stream().synchronize()      # filled before a kernel on another stream reads it
            st = self._state[key] = (t, clear_px)
        self._native.call("rt_push_tiles", dst_ptr, src_ptr, self.width, self.height, clear_px, st[0].data_ptr(), self.bytes.data_ptr(), stream)

    def bytes_moved(self):
        """bytes stored so far (synchronises)"""
        return int(self.bytes.item())


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
        self._copier = None
        self._tiles = None
        self.tile_bytes = None
        if self.ok and self.world > 1:
            torch.cuda.synchronize()
            dist.barrier()      # rank dst's zero-fill of the store (rt_peer_alloc) is complete before anybody writes into it

    def frame(self, k):
        """uint8 view (frame_bytes,) of frame k."""
        return self.memory[k * self.frame_bytes:(k + 1) * self.frame_bytes]

    def frame_ptr(self, k):
        return self._base + k * self.frame_bytes

    def push(self, k, src_ptr, content, stream):
        """Copy-engine push of a locally rendered frame into slot k: only the pixel rect that can differ from what the slot
        holds travels (see SparseFrameCopier; the store starts cleared).  content: the rect Raycaster.render returned.
        Do not mix with kernels rendering into the same slot.  Returns the bytes enqueued."""
        if self._copier is None:
            self._copier = SparseFrameCopier(self.width, self.height)
        return self._copier.copy(k, self._base + k * self.frame_bytes, src_ptr, content, stream)

    def push_tiles(self, k, src_ptr, stream, clear_px=0):
        """Tile-sparse push of a locally rendered frame into slot k (rt_push_tiles): a kernel on THIS GPU stores only the
        32x32-pixel tiles that hold something other than the clear colour -- or did the last time this rank wrote slot k --
        into rank dst's memory.  For frames whose bounding rectangle is most of the frame (a frame-filling raster view) this
        moves about half of what push() moves; seven producers doing that is what keeps rank 0's NVLink ingest below its
        ceiling at N = 8.  Do not mix with push() on the same slot.  The bytes stored accumulate in self.tile_bytes (device)."""
        if self._tiles is None:
            self._tiles = TileFrameCopier(self.width, self.height)
            self.tile_bytes = self._tiles.bytes
        # the store starts zero-filled: exact for clear_px == 0, otherwise every tile has to travel once
        self._tiles.copy(k, self._base + k * self.frame_bytes, src_ptr, stream, clear_px, dirty=clear_px != 0)

    def push_stripes(self, k, src_ptr, content, stripes, stream):
        """Image-space partition: copy-engine push of the row stripes (rows, mod, rem) this rank rendered of frame k, from its
        local full-size frame at src_ptr into slot k -- again only the pixel rect that can differ from what the slot holds
        (cover_rect of the slot's previous content and this frame's; every rank keeps that history for its own stripes).
        Returns the bytes enqueued (an upper bound: whole stripes of the rect)."""
        if self._copier is None:
            self._copier = SparseFrameCopier(self.width, self.height)
        return self._copier.copy_stripes(k, self._base + k * self.frame_bytes, src_ptr, content, stripes, stream)

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
