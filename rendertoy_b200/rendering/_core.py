"""Runtime layer of the `rendering` API on B200: device buffers, images, structs, host matrices.

Mirror of the reference's rendering/_core.py (same public names, argument meaning and error behaviour),
with its pyopencl context/queue/Array/Image objects replaced by torch-owned device memory and the C ABI of
librendertoy_b200.so.  Reference locations are cited per item (paths relative to the reference root).

What changed and why
  * rendering/_core.py:10-17   global cl.Context/CommandQueue/1 GiB zero-filled pool at import
        -> nothing is touched at import; the device is the current torch CUDA device, the stream is
           torch's current stream, the texture pool is allocated on first create_texture2D.
  * :13-14, 323-332            cla.Array            -> DeviceBuffer (same surface tutorials use)
  * :371-373                   cl.Image             -> Image (BGRA8 / float formats in linear memory)
  * :391-418 mapped()          map/unmap            -> copy-out / copy-in; small structs keep a host
                                                      shadow so per-frame `mapped(globals)` costs no
                                                      device round trip
  * :184-310 kernel DSL        OpenCL C text        -> captured verbatim; built-in tutorial shaders are
                                                      recognised by Raster, generic kernels go to
                                                      rendertoy_b200.rendering._dsl (NVRTC)
  * :421-548 host matrices     restated with the NumPy-1.x casting rules they were written for (the
                                                      reference raises under NumPy 2)
"""
import ctypes
import inspect
import math
import os
import typing

import numpy as np
import torch

from .. import _native

# ---------------------------------------------------------------------------------------------------
# vector dtypes (pyopencl.cltypes layout: names s0.., titles x y z w; 3-vectors padded to 4)
# ---------------------------------------------------------------------------------------------------

_C_FLOAT = ctypes.c_float
_ALIGN = {}  # np.dtype -> OpenCL alignment in bytes
_CNAME = {}  # np.dtype -> C type name (for the kernel DSL)


def _vector_dtype(base, count, cname):
    padded = 4 if count == 3 else count
    names = [f"s{i}" for i in range(count)] + [f"padding{i}" for i in range(padded - count)]
    titles = (["x", "y", "z", "w"][:count] + [None] * padded)[:padded]
    fields = [((t, n) if t else n, base) for n, t in zip(names, titles)]
    dt = np.dtype(fields)
    _ALIGN[dt] = dt.itemsize
    _CNAME[dt] = cname
    return dt


float2, float3, float4 = (_vector_dtype(np.float32, n, f"float{n}") for n in (2, 3, 4))
int2, int3, int4 = (_vector_dtype(np.int32, n, f"int{n}") for n in (2, 3, 4))
uint2, uint3, uint4 = (_vector_dtype(np.uint32, n, f"uint{n}") for n in (2, 3, 4))
float4x4 = _vector_dtype(np.float32, 16, "float16")  # rendering/_core.py:112
RGBA = _vector_dtype(np.uint8, 4, "uchar4")          # rendering/_core.py:114

for _s, _n in ((np.float32, "float"), (np.int32, "int"), (np.uint32, "uint"), (np.int64, "long"), (np.uint64, "ulong"),
               (np.uint8, "uchar"), (np.int8, "char"), (np.float64, "double"), (np.int16, "short"), (np.uint16, "ushort")):
    _ALIGN[np.dtype(_s)] = np.dtype(_s).itemsize
    _CNAME[np.dtype(_s)] = _n

r_image1d_t = 'read_only image1d_t'
r_image2d_t = 'read_only image2d_t'
r_image3d_t = 'read_only image3d_t'
w_image1d_t = 'write_only image1d_t'
w_image2d_t = 'write_only image2d_t'
w_image3d_t = 'write_only image3d_t'


def dtype_align(dt):
    dt = np.dtype(dt)
    return _ALIGN.get(dt, dt.alignment)


def dtype_cname(dt):
    return _CNAME[np.dtype(dt)]


# ---------------------------------------------------------------------------------------------------
# constructors (rendering/_core.py:127-181) -- same return types as the reference, including the tuple
# that `.item()` yields when a numpy array is passed to make_float2/4/4x4 and make_int2/4
# ---------------------------------------------------------------------------------------------------

def _make_packed(dt):
    def make(*args):
        if len(args) == 1 and isinstance(args[0], np.ndarray):
            return args[0].ravel().view(dt).item()
        return np.array(args, dtype=dt)
    return make


def _make_padded3(dt, base, zero):
    def make(*args):
        if len(args) == 1 and isinstance(args[0], np.ndarray):
            args = args[0].ravel().view(base)
        return np.array(tuple([*args, zero]), dtype=dt)
    return make


make_int2, make_int4 = _make_packed(int2), _make_packed(int4)
make_float2, make_float4, make_float4x4 = _make_packed(float2), _make_packed(float4), _make_packed(float4x4)
make_int3 = _make_padded3(int3, np.int32, 0)
make_float3 = _make_padded3(float3, np.float32, 0.0)


def to_array(v):
    """Structured vector value(s) -> plain float32 array (rendering/_core.py:169-181)."""
    shape = v.shape
    if shape == ():
        v = v.reshape(1)        # a 0-d structured value cannot be viewed as its scalars
    if v.dtype == float2:
        return v.view(np.float32).reshape(*shape, 2)
    if v.dtype == float3:
        return v.view(np.float32).reshape(*shape, 4)[..., 0:3]
    if v.dtype == float4:
        return v.view(np.float32).reshape(*shape, 4)
    if v.dtype == float4x4:
        return v.view(np.float32).reshape(*shape, 4, 4)
    return v


# ---------------------------------------------------------------------------------------------------
# device memory
# ---------------------------------------------------------------------------------------------------

_HOST_SHADOW_MAX = 64 * 1024


def device():
    """The CUDA device buffers live on: cuda:LOCAL_RANK's current device.  Without CUDA this raises unless
    RENDERTOY_B200_HOST_BUFFERS=1, which places buffers in host memory so host-side logic (loaders, dtype
    layout, matrices) can be exercised; kernels still refuse to run -- there is no CPU compute path."""
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    if os.environ.get("RENDERTOY_B200_HOST_BUFFERS") == "1":
        return torch.device("cpu")
    raise _native.NativeUnavailable(
        "no CUDA device: rendertoy_b200 needs a B200 (set RENDERTOY_B200_HOST_BUFFERS=1 only to test host logic)")


_DEV_INDEX = None


def stream_ptr():
    """cudaStream_t of torch's current stream, passed to every native call.  (torch.cuda.current_stream() costs
    ~15 us of Python per call; the raw getter is ~0.3 us.  One process drives one GPU, so the index is cached.)"""
    global _DEV_INDEX
    if _DEV_INDEX is None:
        _DEV_INDEX = torch.cuda.current_device()
    return torch._C._cuda_getCurrentRawStream(_DEV_INDEX)


class _Storage:
    """One allocation.  `version` bumps on every write so caches (SoA mesh copies, texture objects) know
    when to refresh; allocations <= 64 KiB keep a host shadow."""
    __slots__ = ("tensor", "nbytes", "host", "host_valid", "dev_valid", "version", "cache")

    def __init__(self, nbytes, dev=None, tensor=None):
        self.nbytes = int(nbytes)
        if tensor is not None:      # adopt existing device memory (e.g. a slice of a peer-mapped frame store)
            assert tensor.dtype == torch.uint8 and tensor.is_contiguous() and tensor.numel() >= self.nbytes
            self.tensor = tensor
        else:
            dev = dev or device()
            self.tensor = torch.zeros(max(self.nbytes, 1), dtype=torch.uint8, device=dev)
        # adopted memory has contents (and writers: rt_copy_rect, peer kernels) this object knows nothing about: no shadow
        self.host = np.zeros(self.nbytes, dtype=np.uint8) if (self.nbytes <= _HOST_SHADOW_MAX and tensor is None) else None
        self.host_valid = self.host is not None
        self.dev_valid = True
        self.version = 0
        self.cache = {}

    def sync_device(self):
        if not self.dev_valid:
            self.tensor[:self.nbytes].copy_(torch.from_numpy(self.host))
            self.dev_valid = True

    def read(self, offset, nbytes):
        """Host copy of a byte range (numpy uint8, owned by the caller)."""
        if self.host is not None:
            if not self.host_valid:
                self.host[:] = self.tensor[:self.nbytes].cpu().numpy()
                self.host_valid = True
            return self.host[offset:offset + nbytes].copy()
        return self.tensor[offset:offset + nbytes].cpu().numpy()

    def write(self, offset, data):
        data = np.ascontiguousarray(data).reshape(-1).view(np.uint8)
        if self.host is not None:
            if not self.host_valid:
                self.host[:] = self.tensor[:self.nbytes].cpu().numpy()
                self.host_valid = True
            self.host[offset:offset + data.size] = data
            self.dev_valid = False
        else:
            self.tensor[offset:offset + data.size].copy_(torch.from_numpy(data))
        self.version += 1

    def device_written(self):
        self.host_valid = False
        self.version += 1


class DeviceBuffer:
    """Typed view of device memory with the pyopencl.array.Array surface the tutorials touch
    (rendering/_core.py:13-14, 323-332): shape, dtype, size, nbytes, get(), set(), data, view(), reshape(),
    slicing, len(), map_to_host()."""

    def __init__(self, storage, offset, shape, dtype):
        self._st = storage
        self.offset = int(offset)
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)

    # -- geometry
    @property
    def size(self):
        return int(math.prod(self.shape))

    @property
    def nbytes(self):
        return self.size * self.dtype.itemsize

    @property
    def ndim(self):
        return len(self.shape)

    def __len__(self):
        if not self.shape:
            raise TypeError("len() of a 0-d buffer")
        return self.shape[0]

    # -- device side
    @property
    def ptr(self):
        """Device address (uploads a stale host shadow first)."""
        self._st.sync_device()
        return self._st.tensor.data_ptr() + self.offset

    @property
    def data(self):
        return self

    @property
    def base_data(self):
        return self

    @property
    def version(self):
        return self._st.version

    def tensor(self):
        """The bytes of this view as a torch uint8 tensor (shares memory)."""
        self._st.sync_device()
        return self._st.tensor[self.offset:self.offset + self.nbytes]

    def device_written(self):
        """Call after a kernel wrote through .ptr."""
        self._st.device_written()

    # -- host side
    def get(self):
        raw = self._st.read(self.offset, self.nbytes)
        return raw.view(self.dtype).reshape(self.shape)

    def map_to_host(self):
        return self.get()

    def host_bytes(self):
        """Read-only view of this buffer's bytes on the host without a copy when a valid host shadow exists
        (small structs such as shader globals), else a fresh copy."""
        st = self._st
        if st.host is not None and st.host_valid:
            return st.host[self.offset:self.offset + self.nbytes]
        return st.read(self.offset, self.nbytes)

    def set(self, ary):
        ary = np.asarray(ary)
        if ary.dtype != self.dtype:
            ary = ary.astype(self.dtype)
        if ary.size != self.size:
            raise ValueError(f"cannot set buffer of shape {self.shape} from array of shape {ary.shape}")
        self._st.write(self.offset, ary)

    # -- views
    def view(self, dtype):
        dtype = np.dtype(dtype)
        if not self.shape:
            if dtype.itemsize != self.dtype.itemsize:
                raise ValueError("0-d view must keep the item size")
            return DeviceBuffer(self._st, self.offset, (), dtype)
        last = self.shape[-1] * self.dtype.itemsize
        if last % dtype.itemsize:
            raise ValueError("view: last axis is not a multiple of the new item size")
        return DeviceBuffer(self._st, self.offset, self.shape[:-1] + (last // dtype.itemsize,), dtype)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        if -1 in shape:
            known = -math.prod(shape)
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if math.prod(shape) != self.size:
            raise ValueError(f"cannot reshape {self.shape} into {shape}")
        return DeviceBuffer(self._st, self.offset, shape, self.dtype)

    def ravel(self):
        return self.reshape(self.size)

    def __getitem__(self, idx):
        if not self.shape:
            raise IndexError("0-d buffer cannot be indexed")
        row = math.prod(self.shape[1:]) * self.dtype.itemsize
        if isinstance(idx, slice):
            start, stop, step = idx.indices(self.shape[0])
            if step != 1:
                raise IndexError("only contiguous slices of a device buffer are supported")
            n = max(0, stop - start)
            return DeviceBuffer(self._st, self.offset + start * row, (n,) + self.shape[1:], self.dtype)
        i = int(idx)
        if i < 0:
            i += self.shape[0]
        if not 0 <= i < self.shape[0]:
            raise IndexError("buffer index out of range")
        return DeviceBuffer(self._st, self.offset + i * row, self.shape[1:], self.dtype)

    def __repr__(self):
        return f"DeviceBuffer(shape={self.shape}, dtype={self.dtype.name if self.dtype.names is None else 'struct'}, dev={self._st.tensor.device})"


Buffer = DeviceBuffer  # rendering/_core.py:368


def mesh_soa(vertex_buffer):
    """SoA float4 position / normal arrays of a MeshVertex buffer (rt_mesh_upload_soa), cached on the
    allocation and refreshed when the buffer's version changes.  Returns (pos4, nrm4) torch tensors."""
    st = vertex_buffer._st
    key = ("soa", vertex_buffer.offset, vertex_buffer.size)
    hit = st.cache.get(key)
    if hit is None or hit[0] != st.version:
        assert vertex_buffer.dtype.itemsize == 80, "vertex buffer must hold MeshVertex (80-byte) elements"
        n = vertex_buffer.size
        dev = st.tensor.device
        pos = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
        nrm = torch.empty((max(n, 1), 4), dtype=torch.float32, device=dev)
        _native.call("rt_mesh_upload_soa", vertex_buffer.ptr, n, pos.data_ptr(), nrm.data_ptr(), stream_ptr())
        hit = (st.version, pos, nrm)
        st.cache[key] = hit
    return hit[1], hit[2]


def mesh_bounds(vertex_buffer):
    """Bounding box of a MeshVertex buffer's positions as two (c_double * 3), cached per buffer version like the SoA
    copy.  Costs one small reduction and a device->host read of 24 bytes (a sync) the first time a version is asked for."""
    st = vertex_buffer._st
    key = ("bounds", vertex_buffer.offset, vertex_buffer.size)
    hit = st.cache.get(key)
    if hit is None or hit[0] != st.version:
        pos, _ = mesh_soa(vertex_buffer)
        n = vertex_buffer.size
        if n == 0:
            lo = hi = [float("nan")] * 3
        else:
            xyz = pos[:n, :3]
            both = torch.stack([xyz.amin(0), xyz.amax(0)]).double().cpu().tolist()
            lo, hi = both
        hit = (st.version, (ctypes.c_double * 3)(*lo), (ctypes.c_double * 3)(*hi))
        st.cache[key] = hit
    return hit[1], hit[2]


def mesh_chunk_bounds(vertex_buffer, chunks=64):
    """Bounding boxes of `chunks` runs of consecutive vertices of a MeshVertex buffer as two (c_double * 3k) and k, cached per
    buffer version like mesh_bounds().  The union of their projected rectangles bounds the mesh on screen a quarter tighter
    than the rectangle of its one box (Raster.content_rect); whatever the vertex order, it is a bound."""
    st = vertex_buffer._st
    key = ("chunk_bounds", vertex_buffer.offset, vertex_buffer.size, chunks)
    hit = st.cache.get(key)
    if hit is None or hit[0] != st.version:
        pos, _ = mesh_soa(vertex_buffer)
        n = vertex_buffer.size
        if n == 0:
            lo = hi = np.full((1, 3), np.nan)
        else:
            xyz = pos[:n, :3]
            k = min(chunks, n)
            m = -(-n // k)
            pad = k * m - n
            if pad:
                xyz = torch.cat([xyz, xyz[-1:].expand(pad, 3)])
            both = torch.stack([xyz.view(k, m, 3).amin(1), xyz.view(k, m, 3).amax(1)]).double().cpu().numpy()
            lo, hi = both[0], both[1]
        k = lo.shape[0]
        hit = (st.version, (ctypes.c_double * (3 * k))(*lo.reshape(-1).tolist()), (ctypes.c_double * (3 * k))(*hi.reshape(-1).tolist()), k)
        st.cache[key] = hit
    return hit[1], hit[2], hit[3]


def create_buffer(count: int, dtype: np.dtype):
    """Zero-filled device array (rendering/_core.py:13-14)."""
    dtype = np.dtype(dtype)
    return DeviceBuffer(_Storage(int(count) * dtype.itemsize), 0, (int(count),), dtype)


def create_buffer_from(ary: np.ndarray):
    """rendering/_core.py:323-324"""
    ary = np.ascontiguousarray(ary)
    b = DeviceBuffer(_Storage(ary.nbytes), 0, ary.shape, ary.dtype)
    b.set(ary)
    return b


def create_struct(dtype: np.dtype):
    """0-d zero-filled struct buffer (rendering/_core.py:327-328)."""
    return create_buffer(1, dtype)[0]


def create_struct_from(ary: np.ndarray):
    """rendering/_core.py:331-332"""
    ary = np.asarray(ary)
    b = create_struct(ary.dtype)
    b.set(ary.reshape(()))
    return b


# ---------------------------------------------------------------------------------------------------
# images (rendering/_core.py:335-373)
# ---------------------------------------------------------------------------------------------------

# dtype -> (components, channel numpy dtype as mapped() exposes it, is BGRA8 UNORM)
_IMAGE_FORMATS = {
    float4: (4, np.float32, False),
    float3: (3, np.float32, False),
    float2: (2, np.float32, False),
    np.dtype(np.float32): (1, np.float32, False),
    RGBA: (4, np.int8, True),  # CL_BGRA / CL_UNORM_INT8; the reference maps UNORM_INT8 to np.int8 (:350)
}


def get_valid_image_formats():
    return _IMAGE_FORMATS.keys()


class Image:
    """2-D image in linear device memory, row-major, `components` channels per pixel.  For the RGBA dtype the
    bytes are B,G,R,A (CL_BGRA UNORM8) exactly as the reference's render target (rendering/_core.py:340)."""

    def __init__(self, width, height, dtype, memory=None):
        """memory: optional torch uint8 tensor (width*height*pixel bytes) to use instead of a fresh allocation --
        e.g. one frame of a parallel.FrameStore, which may live on another GPU of the node."""
        dtype = np.dtype(dtype) if not isinstance(dtype, np.dtype) else dtype
        assert dtype in _IMAGE_FORMATS, "Unsupported dtype for image format"
        self.width, self.height, self.depth = int(width), int(height), 0
        self.dtype = dtype
        self.components, self.channel_dtype, self.is_bgra8 = _IMAGE_FORMATS[dtype]
        item = np.dtype(self.channel_dtype).itemsize * self.components
        nbytes = self.width * self.height * item
        self._buffer = DeviceBuffer(_Storage(nbytes, tensor=memory), 0, (self.height, self.width, self.components),
                                    np.dtype(self.channel_dtype))
        self._pending_clear = None   # rgba of a clear() not yet executed (BGRA8 targets only)

    @property
    def shape(self):
        return (self.width, self.height)  # pyopencl Image.shape is (width, height)

    # clear(render_target) is deferred: Raster.draw_triangles folds it into its resolve kernel (no separate fill
    # launch).  Any other access to the pixels executes it first, so the deferral is not observable.
    def flush_clear(self):
        if self._pending_clear is not None:
            rgba, self._pending_clear = self._pending_clear, None
            _native.call("rt_raster_clear_color", self._buffer.ptr, self.width * self.height, rgba, stream_ptr())
            self._buffer.device_written()

    def take_pending_clear(self):
        rgba, self._pending_clear = self._pending_clear, None
        return rgba

    @property
    def buffer(self):
        self.flush_clear()
        return self._buffer

    @property
    def ptr(self):
        return self.buffer.ptr

    @property
    def raw_ptr(self):
        """Device address WITHOUT executing a deferred clear (for callers that consume take_pending_clear())."""
        return self._buffer.ptr

    def get(self):
        """(H, W, C) array; BGRA8 images come back as uint8 bytes B,G,R,A."""
        a = self.buffer.get()
        return a.view(np.uint8) if self.is_bgra8 else a


def create_image2d(width: int, height: int, dtype: np.dtype):
    dtype = np.dtype(dtype)     # the reference's table is keyed by the np.float32 CLASS (:343-349): accept both spellings
    assert dtype in _IMAGE_FORMATS, "Unsupported dtype for image format"
    return Image(width, height, dtype)


# ---------------------------------------------------------------------------------------------------
# clear / mapped (rendering/_core.py:376-418)
# ---------------------------------------------------------------------------------------------------

class DepthView:
    """Raster.get_depth_buffer(): the uint32 depth words that live in the HIGH half of the rasterizer's
    64-bit key buffer.  Behaves like a (W*H,) uint32 buffer for clear(), mapped(), get()."""

    def __init__(self, key_buffer, n_pixels):
        self.key = key_buffer
        self.shape = (int(n_pixels),)
        self.dtype = np.dtype(np.uint32)
        self.size = int(n_pixels)
        self._pending = None   # depth bits of a clear() not yet executed; the next draw (or any read) executes it

    def __len__(self):
        return self.size

    def fill(self, bits, defer=False):
        if defer:
            self._pending = int(bits) & 0xFFFFFFFF
            return
        self._pending = None
        _native.call("rt_raster_clear_depth", self.key.ptr, self.size, int(bits) & 0xFFFFFFFF, stream_ptr())
        self.key.device_written()

    def flush(self):
        if self._pending is not None:
            self.fill(self._pending)

    def take_pending(self):
        bits, self._pending = self._pending, None
        return bits

    def get(self):
        self.flush()
        out = torch.empty(self.size, dtype=torch.int32, device=self.key._st.tensor.device)
        _native.call("rt_raster_read_depth", self.key.ptr, self.size, out.data_ptr(), stream_ptr())
        return out.cpu().numpy().view(np.uint32)

    map_to_host = get

    def set(self, ary):
        self._pending = None
        ary = np.ascontiguousarray(ary, dtype=np.uint32).reshape(-1)
        src = torch.from_numpy(ary.view(np.int32)).to(self.key._st.tensor.device)
        _native.call("rt_raster_write_depth", self.key.ptr, self.size, src.data_ptr(), stream_ptr())
        self.key.device_written()


def clear(b, value=np.float32(0)):
    """Fill a buffer with a repeating value, or an image with a colour (rendering/_core.py:376-388).
    Deviation: for a sub-view the reference fills the whole underlying allocation (b.base_data); this fills
    the view only."""
    if isinstance(b, DepthView) and isinstance(value, (float, np.float32)):      # clear(depth, 1.0): the per-frame call
        b.fill(int(np.float32(value).view(np.uint32)), defer=True)
        return
    if isinstance(b, Image) and b.is_bgra8 and isinstance(value, (float, np.float32)):   # clear(render_target)
        b._pending_clear = _native.float4_const(float(value))
        return
    if isinstance(value, float):
        value = np.float32(value)
    if not isinstance(value, np.ndarray):
        value = np.array(value)
    if isinstance(b, DepthView):
        pat = np.ascontiguousarray(value).reshape(-1).view(np.uint8)
        assert pat.size == 4, "depth buffer is cleared with one 32-bit value"
        b.fill(int(pat.view(np.uint32)[0]), defer=True)
        return
    if isinstance(b, DeviceBuffer):
        pat = np.ascontiguousarray(value).reshape(-1).view(np.uint8)
        assert pat.size and b.nbytes % pat.size == 0, "fill pattern must divide the buffer size"
        b._st.write(b.offset, np.tile(pat, b.nbytes // pat.size))
        return
    assert isinstance(b, Image), "clear() takes a buffer or an image"
    if math.prod(value.shape) <= 1:
        value = np.array([value] * 4)
    rgba = [float(x) for x in np.asarray(value, dtype=np.float32).reshape(-1)[:4]]
    if b.is_bgra8:
        b._pending_clear = _native.float_array(rgba)
    else:
        px = np.asarray(rgba[:b.components], dtype=np.float32)
        b.buffer._st.write(0, np.tile(px, b.width * b.height))


class _Mapped:
    """The context manager mapped() returns (a module-level class: building one per call costs ~7 us of every frame)."""
    __slots__ = ("b", "host", "raw", "direct")

    def __init__(self, b):
        self.b, self.host, self.raw, self.direct = b, None, None, False

    def __enter__(self):
        b = self.b
        self.direct = False
        if isinstance(b, Image):
            a = b.buffer.get()
            self.host = a[..., 0] if b.components == 1 else a
            self.raw = a
        elif isinstance(b, DeviceBuffer) and b._st.host is not None and b._st.host_valid:
            # small buffers (shader globals, descriptors): hand out the host shadow itself, no copies
            self.direct = True
            self.host = b._st.host[b.offset:b.offset + b.nbytes].view(b.dtype).reshape(b.shape)
        else:
            self.host = b.get()
        return self.host

    def __exit__(self, exc_type, exc_val, exc_tb):
        b = self.b
        if isinstance(b, Image):
            b.buffer.set(self.raw)
        elif self.direct:
            b._st.dev_valid = False
            b._st.version += 1
        else:
            b.set(self.host)
        return False


def mapped(b: typing.Union[DeviceBuffer, Image, DepthView]):
    """Context manager giving a writable numpy view of a buffer/image; changes are copied back on exit
    (rendering/_core.py:391-418).  Shapes follow the reference: images map to (H, W, C) (C dropped when 1),
    arrays to their own shape, 0-d structs to a 0-d structured array."""
    return _Mapped(b)


# ---------------------------------------------------------------------------------------------------
# kernel DSL capture (rendering/_core.py:184-310).  The OpenCL C text is kept verbatim; Raster recognises
# the tutorial shader pairs and runs their hand-written CUDA twins, anything else is handed to the NVRTC
# translator in _dsl.py (generic kernels are a "next" row).
# ---------------------------------------------------------------------------------------------------

class _Queue:
    """Placeholder for rendering._core.__queue__/__ctx__ (pyopencl objects in the reference, :10-11)."""

    def finish(self):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()


__ctx__ = _Queue()
__queue__ = __ctx__

_STRUCTS = {}      # C name -> np.dtype, in declaration order
_FUNCTIONS = []    # KernelFunction objects in declaration order


def _get_signature(f):
    signature = inspect.signature(f)
    assert all(v.annotation != inspect.Signature.empty for v in signature.parameters.values()), \
        "All arguments needs to be annotated with a type descriptor"
    return [(k, v) for k, v in signature.parameters.items()], signature.return_annotation


class KernelFunction:
    """What @kernel_function returns (rendering/_core.py:221-228): name, signature, return_annotation; plus
    the captured source."""

    def __init__(self, name, signature, return_annotation, source):
        self.name = name
        self.signature = signature
        self.return_annotation = return_annotation
        self.source = source

    def __call__(self, *args):
        raise Exception("Can not call to this function from host.")


def kernel_function(f):
    s, return_annotation = _get_signature(f)
    k = KernelFunction(f.__name__, s, return_annotation, inspect.getdoc(f) or "")
    _FUNCTIONS.append(k)
    return k


def build_kernel_function(name, arguments, return_type, body):
    class _P:  # minimal stand-in for inspect.Parameter
        def __init__(self, n, a):
            self.name, self.annotation = n, a
    k = KernelFunction(name, [(n, _P(n, a)) for n, a in arguments.items()], return_type, body)
    _FUNCTIONS.append(k)


def build_kernel_main(name, arguments, body):
    from . import _dsl
    return _dsl.Dispatcher(name, dict(arguments), body)


def kernel_main(f):
    s, return_annotation = _get_signature(f)
    assert return_annotation == inspect.Signature.empty, "Kernel main function must return void"
    return build_kernel_main(f.__name__, {v.name: v.annotation for _, v in s}, inspect.getdoc(f) or "")


def kernel_struct(cls):
    """Class annotations -> numpy struct dtype laid out by OpenCL C rules (rendering/_core.py:302-310, where
    pyopencl.tools.match_dtype_to_c_struct does it): every member aligned to its own alignment (vectors to
    their size, float3 to 16), struct aligned and padded to its largest member."""
    fields = cls.__dict__['__annotations__']
    assert all(k in fields.keys() for k in cls.__dict__.keys() if k[0] != "_"), "A public field was declared without annotation"
    names, formats, offsets, off, align = [], [], [], 0, 1
    for k, v in fields.items():
        dt = np.dtype(v)
        a = dtype_align(dt)
        off = (off + a - 1) // a * a
        names.append(k); formats.append(dt); offsets.append(off)
        off += dt.itemsize
        align = max(align, a)
    size = (off + align - 1) // align * align
    dtype = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
    _ALIGN[dtype] = align
    _CNAME[dtype] = cls.__name__
    _STRUCTS[cls.__name__] = dtype
    return dtype


@kernel_struct
class Texture2D:  # rendering/_core.py:313-317
    width: np.int32
    height: np.int32
    offset: np.int32


# ---------------------------------------------------------------------------------------------------
# host transform math (rendering/_core.py:421-548): row-vector convention, left-handed, z in [0,1]
# ---------------------------------------------------------------------------------------------------

def _xyz(args):
    if len(args) == 1:
        return tuple(to_array(args[0]))
    return args


def identity():
    return make_float4x4(*np.eye(4).ravel().tolist())


def translate(*args):
    x, y, z = _xyz(args)
    return make_float4x4(1.0, 0.0, 0.0, 0.0,
                         0.0, 1.0, 0.0, 0.0,
                         0.0, 0.0, 1.0, 0.0,
                         x, y, z, 1.0)


def scale(*args):
    if len(args) == 1 and np.isscalar(args[0]):
        x = y = z = args[0]
    else:
        x, y, z = _xyz(args)
    return make_float4x4(x, 0.0, 0.0, 0.0,
                         0.0, y, 0.0, 0.0,
                         0.0, 0.0, z, 0.0,
                         0.0, 0.0, 0.0, 1.0)


def _r32(x):
    """Round a Python float to float32 and back.  IEEE double has 53 >= 2*24 + 2 mantissa bits, so one double operation on
    float32 operands followed by this rounding IS the float32 operation (double rounding is innocuous for + - * / sqrt):
    the scalar paths below reproduce NumPy's float32 array arithmetic bit for bit without its per-call overhead."""
    return _C_FLOAT(x).value


def _rotate_numpy(angle, axis):
    """rotate() on NumPy scalars: the definition the scalar version is tested against.  Row i, column j of the 3x3 block is
    a_j*a_i*(1-c) plus c on the diagonal or +-a_k*s off it, evaluated left to right with NumPy's own promotion: the axis
    components are float32 scalars, so with a float32 angle everything stays float32, and with a Python / float64 angle the
    products of two axis components are float32 and everything touching cos/sin is float64 -- the same under NumPy 1 and 2,
    and bit-identical to the reference's own rotate() (tests/golden/host_math_reference.npz)."""
    c, s = np.cos(angle), np.sin(angle)
    if not (isinstance(c, np.floating) and c.dtype == np.float32):
        c, s = np.float64(c), np.float64(s)
    a = [np.float32(t) for t in to_array(axis)]
    k = 1 - c
    sign = ((0, 1, -1), (-1, 0, 1), (1, -1, 0))     # sign of a_k * s at (row i, column j), k the third index
    m = []
    for i in range(3):
        for j in range(3):
            term = a[j] * a[i] * k
            m.append(term + c if i == j else (term + a[3 - i - j] * s if sign[i][j] > 0 else term - a[3 - i - j] * s))
        m.append(0)
    return make_float4x4(*[float(t) for t in m], 0.0, 0.0, 0.0, 1.0)


def rotate(angle, axis):
    """Axis-angle rotation with the reference's arithmetic (see _rotate_numpy), on Python scalars with explicit float32
    roundings (see _r32): a tutorial frame calls this once and the NumPy-scalar version costs ~25 us."""
    if not (isinstance(axis, np.ndarray) and axis.dtype == float3 and axis.shape == ()):
        return _rotate_numpy(angle, axis)
    c, s = np.cos(angle), np.sin(angle)
    f32 = isinstance(c, np.floating) and c.dtype == np.float32
    c, s = float(c), float(s)
    x, y, z, _ = axis.item()
    xx, yy, zz, xy, xz, yz = _r32(x * x), _r32(y * y), _r32(z * z), _r32(x * y), _r32(x * z), _r32(y * z)
    if f32:     # float32 angle: every operation rounds to float32
        k = _r32(1 - c)
        xs, ys, zs = _r32(x * s), _r32(y * s), _r32(z * s)
        r = _r32
        return make_float4x4(r(r(xx * k) + c), r(r(xy * k) + zs), r(r(xz * k) - ys), 0.0,
                             r(r(xy * k) - zs), r(r(yy * k) + c), r(r(yz * k) + xs), 0.0,
                             r(r(xz * k) + ys), r(r(yz * k) - xs), r(r(zz * k) + c), 0.0,
                             0.0, 0.0, 0.0, 1.0)
    k = 1 - c
    return make_float4x4(xx * k + c, xy * k + z * s, xz * k - y * s, 0.0,
                         xy * k - z * s, yy * k + c, yz * k + x * s, 0.0,
                         xz * k + y * s, yz * k - x * s, zz * k + c, 0.0,
                         0.0, 0.0, 0.0, 1.0)


def matmul(a, b):
    assert a.dtype == float4 or a.dtype == float4x4, "First vector must be a float4 or a matrix float4x4"
    assert b.dtype == float4x4, "Second argument must be a matrix"
    a_is_vec = a.dtype == float4
    c = to_array(a) @ to_array(b)
    return make_float4(c) if a_is_vec else make_float4x4(c)


def dot(v1, v2):
    assert v1.dtype == v2.dtype, "Can not apply dot product between different vector types"
    assert v1.shape == v2.shape, "Can not apply dot product between different vector types"
    if v1.dtype not in (float2, float3, float4):
        raise Exception('Not valid dtype')
    n = {float2: 2, float3: 3, float4: 4}[v1.dtype]
    acc = v1['x'] * v2['x']
    for c in "yzw"[:n - 1]:
        acc = acc + v1[c] * v2[c]          # float32 products and sums, left to right
    return acc.item()


def normalize(v):
    v_dtype = v.dtype
    l = np.float32(np.sqrt(dot(v, v)))       # NumPy 1: float32 array / float64 scalar stays float32
    if v_dtype == float3:
        return make_float3((to_array(v) / l).astype(np.float32))
    # the reference divides twice on this branch (:510); kept
    return ((to_array(v) / l).astype(np.float32) / l).astype(np.float32).view(v_dtype)


def cross(v1, v2):
    return make_float3(
        (v1['y'] * v2['z'] - v1['z'] * v2['y']).item(),
        (v1['z'] * v2['x'] - v1['x'] * v2['z']).item(),
        (v1['x'] * v2['y'] - v1['y'] * v2['x']).item()
    )


def direction(f, t):
    return normalize(make_float3((to_array(t) - to_array(f)).astype(np.float32)))


def _look_at_numpy(camera, target, up_vector):
    """look_at() through the public vector helpers; kept as the definition the scalar version is tested against."""
    zaxis = direction(camera, target)
    xaxis = normalize(cross(up_vector, zaxis))
    yaxis = cross(zaxis, xaxis)
    cols = [xaxis, yaxis, zaxis]
    rows = [[c[k] for c in cols] + [0] for k in "xyz"]
    rows.append([-dot(c, camera) for c in cols] + [1])
    return make_float4x4(*[e for r in rows for e in r])


def _normalize3(x, y, z):
    """normalize() of a float3 on scalars: float32 dot, float64 sqrt rounded to float32, float32 divisions."""
    l = _r32(math.sqrt(_r32(_r32(_r32(x * x) + _r32(y * y)) + _r32(z * z))))
    if l == 0.0:
        return None
    return _r32(x / l), _r32(y / l), _r32(z / l)


def _cross3(a, b):
    return (_r32(_r32(a[1] * b[2]) - _r32(a[2] * b[1])), _r32(_r32(a[2] * b[0]) - _r32(a[0] * b[2])),
            _r32(_r32(a[0] * b[1]) - _r32(a[1] * b[0])))


def _dot3(a, b):
    return _r32(_r32(_r32(a[0] * b[0]) + _r32(a[1] * b[1])) + _r32(a[2] * b[2]))


def look_at(camera, target, up_vector):
    """View matrix with the camera axes as columns (rendering/_core.py:528-538).  Scalar float32 arithmetic (see _r32),
    bit-identical to the helper-based formulation in _look_at_numpy, which costs ~60 us per call."""
    ok = all(isinstance(v, np.ndarray) and v.dtype == float3 and v.shape == () for v in (camera, target, up_vector))
    if not ok:
        return _look_at_numpy(camera, target, up_vector)
    cam, tgt, up = camera.item()[:3], target.item()[:3], up_vector.item()[:3]
    zaxis = _normalize3(_r32(tgt[0] - cam[0]), _r32(tgt[1] - cam[1]), _r32(tgt[2] - cam[2]))
    xaxis = _normalize3(*_cross3(up, zaxis)) if zaxis is not None else None
    if xaxis is None:       # degenerate input: let NumPy produce its inf/nan pattern (and warnings) as before
        return _look_at_numpy(camera, target, up_vector)
    yaxis = _cross3(zaxis, xaxis)
    return make_float4x4(xaxis[0], yaxis[0], zaxis[0], 0,
                         xaxis[1], yaxis[1], zaxis[1], 0,
                         xaxis[2], yaxis[2], zaxis[2], 0,
                         -_dot3(xaxis, cam), -_dot3(yaxis, cam), -_dot3(zaxis, cam), 1)


def perspective(fov=3.141593 / 4, aspect_ratio=1.0, znear=.01, zfar=100.0):
    hs = 1.0 / np.tan(fov / 2)
    ws = hs / aspect_ratio
    return make_float4x4(ws, 0, 0, 0,
                         0, hs, 0, 0,
                         0, 0, zfar / (zfar - znear), 1.0,
                         0, 0, -znear * zfar / (zfar - znear), 0)


# ---------------------------------------------------------------------------------------------------
# textures (rendering/_core.py:551-578): bump-allocated float4 textures in one pool + Texture2D descriptor.
# sample2D's raw-pointer gather becomes a point-sampled CUDA texture object per allocation.
# ---------------------------------------------------------------------------------------------------

__MAX_SIZE__ = int(os.environ.get("RENDERTOY_B200_POOL_BYTES", 1024 * 1024 * 1024))
_TEX_ALIGN = 512  # cudaDeviceProp::textureAlignment for linear-memory texture objects


class MemoryPool:
    def __init__(self):
        self.max_size = __MAX_SIZE__
        self.buffer = None  # allocated on first use (the reference zero-fills 1 GiB at import, :16-17)
        self.malloc_ptr = 0
        self.textures = {}  # offset -> dict(width, height, view, handle)

    def get_buffer(self):
        if self.buffer is None:
            self.buffer = create_buffer(self.max_size, np.uint8)
        return self.buffer

    def allocate_texture(self, width, height):
        memory_to_allocate = width * height * 4 * 4
        if self.malloc_ptr + memory_to_allocate >= self.max_size:
            raise Exception("Memory out!")
        memory = self.get_buffer()[self.malloc_ptr:self.malloc_ptr + memory_to_allocate]
        view = memory.view(float4).reshape(height, width)
        texture_descriptor = create_struct(Texture2D)
        with mapped(texture_descriptor) as map:
            map['width'] = width
            map['height'] = height
            map['offset'] = self.malloc_ptr
        self.textures[self.malloc_ptr] = {"width": width, "height": height, "view": view, "handle": 0}
        self.malloc_ptr += (memory_to_allocate + _TEX_ALIGN - 1) // _TEX_ALIGN * _TEX_ALIGN
        return view, texture_descriptor

    def texture_handle(self, offset):
        """Native texture object for the allocation starting at `offset` (created on first use)."""
        t = self.textures[int(offset)]
        if not t["handle"]:
            import ctypes
            h = ctypes.c_uint64(0)
            _native.call("rt_texture_create", t["view"].ptr, t["width"], t["height"], ctypes.byref(h))
            t["handle"] = h.value
        else:
            t["view"]._st.sync_device()
        return t["handle"]


__MEMORY_POOL__ = MemoryPool()


def create_texture2D(width: int, height: int):
    return __MEMORY_POOL__.allocate_texture(width, height)
