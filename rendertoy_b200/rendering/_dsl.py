"""Generic kernel dispatch: @kernel_main / build_kernel_main on NVRTC (reference: rendering/_core.py:247-299).

The reference appends every @kernel_struct, @kernel_function and @kernel_main to one OpenCL C string and builds it
at the first dispatch (`cl.Program(__ctx__, __code__).build()`, :283-284).  Same model here: the captured OpenCL C
is assembled into ONE CUDA C++ translation unit over cl_prelude.cuh (a handful of textual rewrites, listed in
to_cuda()), compiled by NVRTC for sm_100a through the C ABI (rt_dsl_compile) and launched 1-D with the reference's
conventions: `int thread_id`, the `number_of_threads` guard (:252-253), pointer arguments for `[T]` annotations,
by-value copies for struct / scalar annotations (:269-279), None -> NULL.

The two raster hot-path shader pairs do NOT go through here (they have hand-written kernels, see _raster.py); this is
the SURVEY.md section 8f.1 row: lessons 01-07 style compute kernels.  There is no CPU fallback: without NVRTC or a GPU
the dispatch raises.
"""
import ctypes
import hashlib
import math
import os
import re

import numpy as np

from .. import _native
from . import _core

_HERE = os.path.dirname(os.path.abspath(__file__))
_VEC_TYPES = "float2|float3|float4|float8|float16|float4x4|int2|int3|int4"
_MODULES = {}      # source hash -> native module handle
KERNELS = []       # Dispatcher objects in declaration order


class _ClImageArg(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("width", ctypes.c_int), ("height", ctypes.c_int), ("components", ctypes.c_int),
                ("is_unorm8_bgra", ctypes.c_int)]


def _ctype(annotation):
    """C spelling of an annotation (rendering/_core.py:198-208)."""
    if annotation is None:
        return "void"
    ptr = False
    if isinstance(annotation, list):
        assert len(annotation) == 1, "parameters annotated with list should refer to a pointer to a single type, e.g. [int] is considered a int*."
        ptr, annotation = True, annotation[0]
    if isinstance(annotation, str):      # 'write_only image2d_t' and friends
        return annotation
    return _core.dtype_cname(annotation) + ("*" if ptr else "")


def to_cuda(opencl_c: str) -> str:
    """OpenCL C -> CUDA C++ over cl_prelude.cuh: vector literals `(float4)(...)` become constructor calls, the
    two-level float16 swizzles of mul() get one-level names.  Everything else is handled by the prelude's types,
    operators and #defines."""
    s = re.sub(r"\((%s)\)\s*\(" % _VEC_TYPES, lambda m: "cl_mk_%s(" % ("float16" if m.group(1) == "float4x4" else m.group(1)), opencl_c)
    for a, b in ((".even.even", ".even_even_"), (".odd.even", ".odd_even_"), (".even.odd", ".even_odd_"), (".odd.odd", ".odd_odd_")):
        s = s.replace(a, b)
    return s


def _reachable_functions(body):
    """@kernel_function definitions a kernel body uses, transitively, in declaration order (latest definition of a
    name wins: interactive sessions and test suites re-declare shaders; a script declares each once)."""
    latest = {}
    for f in _core._FUNCTIONS:
        latest[f.name] = f
    used, frontier = set(), [body]
    while frontier:
        text = frontier.pop()
        for name, f in latest.items():
            if name not in used and re.search(r"\b%s\s*\(" % re.escape(name), text):
                used.add(name)
                frontier.append(f.source)
    return [f for f in latest.values() if f.name in used]


def program_source(kernel):
    """One kernel's program: prelude, every declared struct, the device functions it reaches, the kernel itself
    (the reference compiles its whole accumulated __code__ instead, :283-284; the observable behaviour is the same)."""
    parts = [open(os.path.join(_HERE, "cl_prelude.cuh")).read()]
    pool = _core.__MEMORY_POOL__
    functions = _reachable_functions(kernel.body)
    uses_textures = any("sample2D" in f.source for f in functions) or "sample2D" in kernel.body
    if uses_textures:
        parts.append(f"__device__ const unsigned long long memory_pool_ptr = {pool.get_buffer().ptr}ull;\n"
                     "#define sample2D(texture, c) cl_sample2D(memory_pool_ptr, (texture), (c))\n"
                     "#define sample2D_linear(texture, c) cl_sample2D_linear(memory_pool_ptr, (texture), (c))\n")
    for name, dt in _core._STRUCTS.items():
        fields = "\n".join(f"    {_core.dtype_cname(dt.fields[n][0])} {n};" for n in dt.names)
        parts.append(f"struct {name} {{\n{fields}\n}};\nstatic_assert(sizeof({name}) == {dt.itemsize}, \"{name}: layout differs from the host dtype\");\n")
    for f in functions:
        sig = ", ".join(f"{_ctype(p.annotation)} {p.name}" for _, p in f.signature)
        parts.append(f"__device__ {_ctype(f.return_annotation)} {f.name}({sig}) {{\n{to_cuda(f.source)}\n}}\n")
    for k in (kernel,):
        sig = ", ".join([f"{_ctype(a)} {n}" for n, a in k.arguments.items()] + ["int number_of_threads"])
        parts.append(f"extern \"C\" __global__ void {k.name}({sig}) {{\n    int thread_id = get_global_id(0);\n"
                     f"    if (thread_id >= number_of_threads) return; // automatically skip threads outside range\n{to_cuda(k.body)}\n}}\n")
    return "\n".join(parts)


def compile_program(source):
    """NVRTC-compile (cached by content).  Works without a GPU; loading/launching needs one."""
    key = hashlib.sha256(source.encode()).hexdigest()
    if key not in _MODULES:
        handle = ctypes.c_uint64(0)
        log = ctypes.create_string_buffer(1 << 16)
        rc = _native.lib().rt_dsl_compile(source.encode(), ctypes.byref(handle), log, len(log))
        if rc != 0:
            raise RuntimeError(f"kernel program failed to build: {_native.lib().rt_last_error().decode()}\n{log.value.decode(errors='replace')}")
        _MODULES[key] = handle.value
    return _MODULES[key]


def launch(module, kernel_name, n_threads, values):
    """Launch a kernel of a compiled module.  values: ctypes objects (pointers / structures) or numpy values / arrays
    (passed by value, byte for byte); the trailing `int number_of_threads` is appended here."""
    keep, slots = [], (ctypes.c_void_p * (len(values) + 1))()
    for i, v in enumerate(values):
        if not isinstance(v, (ctypes._SimpleCData, ctypes.Structure, ctypes.Array)):
            raw = np.ascontiguousarray(v).reshape(-1).view(np.uint8).copy()
            keep.append(raw)
            v = (ctypes.c_uint8 * max(raw.size, 1)).from_buffer(raw)
        keep.append(v)
        slots[i] = ctypes.cast(ctypes.pointer(v), ctypes.c_void_p)
    n = ctypes.c_int(int(n_threads))
    slots[len(values)] = ctypes.cast(ctypes.pointer(n), ctypes.c_void_p)
    _native.call("rt_dsl_launch", module, kernel_name.encode(), int(n_threads), slots, _core.stream_ptr())


def raster_program_source(vertex_shader, fragment_shader, vin, vout, vsg, fsg):
    """Program for Raster with user shaders: prelude, structs, the shaders and what they call, raster_generic.cuh."""
    parts = [open(os.path.join(_HERE, "cl_prelude.cuh")).read()]
    functions = _reachable_functions(f"{vertex_shader.name}( {fragment_shader.name}(")
    if any("sample2D" in f.source for f in functions):
        parts.append(f"__device__ const unsigned long long memory_pool_ptr = {_core.__MEMORY_POOL__.get_buffer().ptr}ull;\n"
                     "#define sample2D(texture, c) cl_sample2D(memory_pool_ptr, (texture), (c))\n"
                     "#define sample2D_linear(texture, c) cl_sample2D_linear(memory_pool_ptr, (texture), (c))\n")
    for name, dt in _core._STRUCTS.items():
        fields = "\n".join(f"    {_core.dtype_cname(dt.fields[n][0])} {n};" for n in dt.names)
        parts.append(f"struct {name} {{\n{fields}\n}};\nstatic_assert(sizeof({name}) == {dt.itemsize}, \"{name}: layout differs from the host dtype\");\n")
    for f in functions:
        sig = ", ".join(f"{_ctype(p.annotation)} {p.name}" for _, p in f.signature)
        parts.append(f"__device__ {_ctype(f.return_annotation)} {f.name}({sig}) {{\n{to_cuda(f.source)}\n}}\n")
    parts.append(f"#define VIN_T {_ctype(vin)}\n#define VOUT_T {_ctype(vout)}\n#define VSG_T {_ctype(vsg)}\n#define FSG_T {_ctype(fsg)}\n"
                 f"#define VS_FN {vertex_shader.name}\n#define FS_FN {fragment_shader.name}\n")
    parts.append(open(os.path.join(_HERE, "raster_generic.cuh")).read())
    return "\n".join(parts)


class Dispatcher:
    """`kernel[n](*args)` (rendering/_core.py:261-290)."""

    def __init__(self, name, arguments, body):
        self.name, self.arguments, self.body = name, arguments, body
        KERNELS.append(self)

    def __getitem__(self, num_threads):
        if isinstance(num_threads, (list, tuple)):
            num_threads = math.prod(num_threads)

        def dispatch_call(*args):
            module = compile_program(program_source(self))   # built lazily, like :283-284
            keep, slots, written = [], (ctypes.c_void_p * (len(self.arguments) + 1))(), []
            assert len(args) == len(self.arguments), f"{self.name} takes {len(self.arguments)} arguments"
            for i, (a, (pname, ann)) in enumerate(zip(args, self.arguments.items())):
                if isinstance(a, (_core.DeviceBuffer, _core.DepthView)) and isinstance(ann, list):
                    assert isinstance(a, _core.DeviceBuffer), "the depth buffer view cannot be bound to a kernel pointer"
                    v = ctypes.c_void_p(a.ptr)
                    written.append(a)
                elif isinstance(a, _core.Image):
                    v = _ClImageArg(a.ptr, a.width, a.height, a.components, int(a.is_bgra8))
                    written.append(a.buffer)
                elif a is None:
                    v = ctypes.c_void_p(0)
                else:
                    if isinstance(a, _core.DeviceBuffer):
                        a = a.get()                       # value annotation: copied to the host, passed by value (:274)
                    if isinstance(a, int):
                        a = np.int32(a)
                    if isinstance(a, float):
                        a = np.float32(a)
                    raw = np.ascontiguousarray(a).reshape(-1).view(np.uint8).copy()
                    v = (ctypes.c_uint8 * max(raw.size, 1)).from_buffer(raw)
                    keep.append(raw)
                keep.append(v)
                slots[i] = ctypes.cast(ctypes.pointer(v), ctypes.c_void_p)
            n = ctypes.c_int(int(num_threads))
            keep.append(n)
            slots[len(self.arguments)] = ctypes.cast(ctypes.pointer(n), ctypes.c_void_p)
            _native.call("rt_dsl_launch", module, self.name.encode(), int(num_threads), slots, _core.stream_ptr())
            for b in written:
                b.device_written()

        return dispatch_call
