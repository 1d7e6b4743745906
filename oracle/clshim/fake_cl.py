"""oracle/clshim/fake_cl.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU stand-in for the slice of pyopencl (+ stubs for pywavefront / sdl2) that the reference package imports,
so that the UNMODIFIED reference Python (`/root/reference/rendering`) and its OWN OpenCL C kernel strings can
run in this container, which has no pyopencl wheel, no OpenCL ICD and no network.  Kernels are compiled for
the host with g++ through cl_compat.hpp (see translate.py) and executed one work-item after another, in
global-id order, on host memory.

    import oracle.clshim.fake_cl as fake_cl; fake_cl.install(); sys.path.insert(0, "/root/reference")
    import rendering            # the reference itself

Used only by oracle/clshim/make_golden.py to pin oracle/raster_oracle.c (tests/golden/*.npz).
"""
import ctypes
import enum
import sys
import types

import numpy as np

from . import translate

# ---------------------------------------------------------------------------------------------------------
# dtype registry (pyopencl.tools / pyopencl.cltypes)
# ---------------------------------------------------------------------------------------------------------
_DTYPE_BY_NAME = {}
_NAME_BY_DTYPE = {}
_ALIGN = {}


def _register(name, dtype, align=None):
    dtype = np.dtype(dtype)
    _DTYPE_BY_NAME[name] = dtype
    _NAME_BY_DTYPE.setdefault(dtype, name)
    _ALIGN[dtype] = align or dtype.itemsize


for _np, _c in ((np.float32, "float"), (np.int32, "int"), (np.uint32, "uint"), (np.int64, "long"), (np.uint64, "ulong"),
                (np.uint8, "uchar"), (np.int8, "char"), (np.float64, "double")):
    _register(_c, _np)


def _vector(base, base_name, count):
    padded = 4 if count == 3 else count
    names = [f"s{i}" for i in range(count)] + [f"padding{i}" for i in range(padded - count)]
    titles = (["x", "y", "z", "w"][:count] + [None] * padded)[:padded]
    dt = np.dtype([((t, n) if t else n, base) for n, t in zip(names, titles)])
    _register(f"{base_name}{count}", dt)


for _b, _n in ((np.float32, "float"), (np.int32, "int"), (np.uint32, "uint"), (np.uint8, "uchar")):
    for _cnt in (2, 3, 4, 8, 16):
        _vector(_b, _n, _cnt)


def get_or_register_dtype(name, dtype=None):
    if dtype is not None:
        _register(name, dtype, _ALIGN.get(np.dtype(dtype)))
        return np.dtype(dtype)
    return _DTYPE_BY_NAME[name]


def dtype_to_ctype(dtype):
    return _NAME_BY_DTYPE[np.dtype(dtype)]


def match_dtype_to_c_struct(device, name, dtype):
    """OpenCL C layout of a struct + its C declaration (pyopencl.tools.match_dtype_to_c_struct)."""
    names, formats, offsets, off, align = [], [], [], 0, 1
    lines = []
    for fname in dtype.names:
        ft = dtype.fields[fname][0]
        a = _ALIGN.get(ft, ft.alignment)
        off = (off + a - 1) // a * a
        names.append(fname); formats.append(ft); offsets.append(off)
        off += ft.itemsize
        align = max(align, a)
        lines.append(f"  {dtype_to_ctype(ft)} {fname};")
    size = (off + align - 1) // align * align
    out = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": size})
    _ALIGN[out] = align
    decl = "typedef struct {\n" + "\n".join(lines) + f"\n}} {name};\nstatic_assert(sizeof({name}) == {size}, \"layout\");\n"
    return out, decl


# ---------------------------------------------------------------------------------------------------------
# memory objects
# ---------------------------------------------------------------------------------------------------------
class _Mapped(np.ndarray):
    def release(self):
        pass


class Buffer:
    """Raw device allocation == a host byte array."""

    def __init__(self, nbytes=0, host=None, offset=0):
        self.host = host if host is not None else np.zeros(max(int(nbytes), 1), dtype=np.uint8)
        self.offset = offset
        self.size = int(nbytes)

    @property
    def address(self):
        return self.host.ctypes.data + self.offset


class Array:
    """pyopencl.array.Array on host memory."""

    def __init__(self, queue, shape, dtype, base=None, offset=0):
        self.queue = queue
        self.shape = tuple(shape) if isinstance(shape, (tuple, list)) else (int(shape),)
        self.dtype = np.dtype(dtype)
        self.size = int(np.prod(self.shape)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self.base_data = base if base is not None else Buffer(self.nbytes)
        self.offset = offset

    @property
    def data(self):
        return Buffer(self.nbytes, self.base_data.host, self.base_data.offset + self.offset)

    def _np(self):
        raw = self.base_data.host[self.base_data.offset + self.offset:self.base_data.offset + self.offset + self.nbytes]
        return raw.view(self.dtype).reshape(self.shape)

    def get(self):
        return self._np().copy()

    def map_to_host(self):
        return self._np()

    def set(self, ary):
        self._np()[...] = ary

    def view(self, dtype):
        dtype = np.dtype(dtype)
        last = self.shape[-1] * self.dtype.itemsize // dtype.itemsize
        return Array(self.queue, self.shape[:-1] + (last,), dtype, self.base_data, self.offset)

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list)):
            shape = tuple(shape[0])
        return Array(self.queue, shape, self.dtype, self.base_data, self.offset)

    def __len__(self):
        return self.shape[0]

    def __getitem__(self, idx):
        row = int(np.prod(self.shape[1:])) * self.dtype.itemsize
        if isinstance(idx, slice):
            start, stop, _ = idx.indices(self.shape[0])
            return Array(self.queue, (max(0, stop - start),) + self.shape[1:], self.dtype, self.base_data, self.offset + start * row)
        return Array(self.queue, self.shape[1:], self.dtype, self.base_data, self.offset + int(idx) * row)


class _ClImageStruct(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_int), ("height", ctypes.c_int), ("components", ctypes.c_int),
                ("is_unorm8_bgra", ctypes.c_int)]


class channel_order(enum.IntEnum):
    R = 0; RG = 1; RGB = 2; RGBA = 3; BGRA = 4


class channel_type(enum.IntEnum):
    FLOAT = 0; SIGNED_INT32 = 1; SIGNED_INT8 = 2; UNSIGNED_INT32 = 3; UNSIGNED_INT8 = 4; UNORM_INT8 = 5


class mem_object_type(enum.IntEnum):
    BUFFER = 0; IMAGE1D = 1; IMAGE2D = 2; IMAGE3D = 3


class mem_flags(enum.IntFlag):
    READ_WRITE = 1; READ_ONLY = 2; WRITE_ONLY = 4


class map_flags(enum.IntFlag):
    READ = 1; WRITE = 2


class device_info(enum.IntEnum):
    PREFERRED_WORK_GROUP_SIZE_MULTIPLE = 0


class ImageFormat:
    def __init__(self, order, ctype):
        self.channel_order, self.channel_data_type = order, ctype


_COMPONENTS = {channel_order.R: 1, channel_order.RG: 2, channel_order.RGB: 3, channel_order.RGBA: 4, channel_order.BGRA: 4}
_CH_BYTES = {channel_type.FLOAT: 4, channel_type.SIGNED_INT32: 4, channel_type.UNSIGNED_INT32: 4, channel_type.SIGNED_INT8: 1,
             channel_type.UNSIGNED_INT8: 1, channel_type.UNORM_INT8: 1}


class Image:
    def __init__(self, ctx, flags, fmt, shape):
        self.format = fmt
        self.width, self.height = int(shape[0]), int(shape[1]) if len(shape) > 1 else 0
        self.depth = 0
        self.shape = tuple(shape)
        self.type = mem_object_type.IMAGE2D
        self.components = _COMPONENTS[fmt.channel_order]
        self.host = np.zeros(self.width * max(1, self.height) * self.components * _CH_BYTES[fmt.channel_data_type], dtype=np.uint8)
        self.cstruct = _ClImageStruct(self.host.ctypes.data, self.width, self.height, self.components,
                                      int(fmt.channel_order == channel_order.BGRA and fmt.channel_data_type == channel_type.UNORM_INT8))


class Context:
    devices = [object()]


def create_some_context():
    return Context()


class CommandQueue:
    def __init__(self, ctx):
        self.ctx = ctx


def zeros(queue, shape, dtype):
    return Array(queue, shape, dtype)


def to_device(queue, ary):
    ary = np.asarray(ary)
    a = Array(queue, ary.shape, ary.dtype)
    a.set(ary)
    return a


def enqueue_fill_buffer(queue, buf, pattern, offset, size):
    pat = np.ascontiguousarray(pattern).reshape(-1).view(np.uint8)
    raw = buf.host[buf.offset + offset:buf.offset + offset + size]
    raw[:] = np.tile(pat, size // pat.size)


def enqueue_fill_image(queue, img, color, origin, region):
    color = np.asarray(color, dtype=np.float32).reshape(-1)
    if img.cstruct.is_unorm8_bgra:
        v = np.clip(color[[2, 1, 0, 3]] * np.float32(255.0), 0, 255)
        px = np.rint(v).astype(np.uint8)
        img.host.reshape(-1, 4)[:] = px
    else:
        img.host.view(np.float32).reshape(-1, img.components)[:] = color[:img.components]


class _Event:
    def wait(self):
        pass


def enqueue_map_buffer(queue, buf, flags, offset, shape, dtype):
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if shape else 1
    raw = buf.host[buf.offset + offset:buf.offset + offset + n * dtype.itemsize].view(_Mapped)
    return raw.view(dtype).reshape(shape), _Event()


def enqueue_map_image(queue, img, flags, origin, region, shape, dtype):
    return img.host.view(_Mapped).view(np.dtype(dtype)).reshape(shape), _Event()


class _Kernel:
    def __init__(self, program, name):
        self.program, self.name = program, name
        self.fn = getattr(program.lib, name + "__launch")
        self.fn.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_long]
        self.fn.restype = None
        self.params = program.kernels[name]

    def __call__(self, queue, global_size, local_size, *args):
        assert len(args) == len(self.params), f"{self.name}: expected {len(self.params)} args, got {len(args)}"
        keep, slots = [], (ctypes.c_void_p * len(args))()
        for i, (a, (ctype, is_ptr)) in enumerate(zip(args, self.params)):
            if is_ptr or ctype == "image2d_t":
                if a is None:
                    slots[i] = None
                elif isinstance(a, Buffer):
                    slots[i] = a.address
                elif isinstance(a, Image):
                    slots[i] = ctypes.addressof(a.cstruct)
                else:
                    raise TypeError(f"{self.name} arg {i}: cannot pass {type(a)} as a pointer")
            else:
                v = np.ascontiguousarray(a).reshape(-1).view(np.uint8).copy()
                keep.append(v)
                slots[i] = v.ctypes.data
        self.fn(slots, int(np.prod(global_size)))


class Program:
    def __init__(self, ctx, source):
        self.source = source

    def build(self):
        self.lib, self.kernels = translate.compile_program(self.source)
        return self

    def __getattr__(self, name):
        if name.startswith("_") or name in ("lib", "kernels", "source"):
            raise AttributeError(name)
        return _Kernel(self, name)


def install():
    """Put the fake modules into sys.modules so `import pyopencl` etc. in the reference resolve to this file."""
    cl = types.ModuleType("pyopencl")
    for k in ("Buffer", "Image", "ImageFormat", "Context", "CommandQueue", "Program", "create_some_context", "channel_order",
              "channel_type", "mem_object_type", "mem_flags", "map_flags", "device_info", "enqueue_fill_buffer", "enqueue_fill_image",
              "enqueue_map_buffer", "enqueue_map_image"):
        setattr(cl, k, globals()[k])
    cla = types.ModuleType("pyopencl.array")
    cla.Array, cla.zeros, cla.to_device = Array, zeros, to_device
    tools = types.ModuleType("pyopencl.tools")
    tools.get_or_register_dtype, tools.dtype_to_ctype, tools.match_dtype_to_c_struct = get_or_register_dtype, dtype_to_ctype, match_dtype_to_c_struct
    cl.array, cl.tools = cla, tools
    sys.modules.update({"pyopencl": cl, "pyopencl.array": cla, "pyopencl.tools": tools})
    # display / OBJ parsing are not on the path being pinned: inert stubs
    pw = types.ModuleType("pywavefront")
    pw.Wavefront = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("pywavefront is not available in the shim"))
    sdl2 = types.ModuleType("sdl2")
    sdl2.ext = types.ModuleType("sdl2.ext")
    sys.modules.update({"pywavefront": pw, "sdl2": sdl2, "sdl2.ext": sdl2.ext})
