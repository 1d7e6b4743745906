"""ctypes binding of librendertoy_b200.so (the C ABI declared in include/rendertoy_b200.h).

The library is the product's only compute path.  If it is missing or fails to load, every call raises
NativeUnavailable -- there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RENDERTOY_B200_LIB") or os.path.join(HERE, "librendertoy_b200.so")   # override: A/B builds

BVH_LBVH, BVH_PLOC = 0, 1
SHADER_LESSON08 = 8
SHADER_LESSON09 = 9
NO_PRIMITIVE = 0xFFFFFFFF

# name -> (restype, argtypes); mirrors include/rendertoy_b200.h one to one
_VP, _I64, _I32, _U64, _U32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint64, C.c_uint32
_FP = C.POINTER(C.c_float)
SIGNATURES = {
    "rt_abi_version": (C.c_int, []),
    "rt_last_error": (C.c_char_p, []),
    "rt_device_info": (C.c_int, [C.POINTER(C.c_int)] * 4),
    "rt_mesh_upload_soa": (C.c_int, [_VP, _I64, _VP, _VP, _VP]),
    "rt_raster_clear_depth": (C.c_int, [_VP, _I64, _U32, _VP]),
    "rt_raster_clear_color": (C.c_int, [_VP, _I64, _FP, _VP]),
    "rt_raster_read_depth": (C.c_int, [_VP, _I64, _VP, _VP]),
    "rt_raster_write_depth": (C.c_int, [_VP, _I64, _VP, _VP]),
    "rt_raster_scratch_bytes": (_I64, [_I32, _I64, _I32, _I32]),
    "rt_raster_draw_triangles": (C.c_int, [_VP, _VP, _VP, _I64, _I32, _FP, _U64, _I32, _I32, _VP, _VP, _I64, _VP, _FP, _I32, _U32, C.POINTER(C.c_int), _VP]),
    "rt_raster_screen_bounds": (C.c_int, [_FP, C.POINTER(C.c_double), C.POINTER(C.c_double), _I32, _I32, C.POINTER(C.c_int)]),
    "rt_raster_screen_bounds_n": (C.c_int, [_FP, C.POINTER(C.c_double), C.POINTER(C.c_double), _I32, _I32, _I32, C.POINTER(C.c_int)]),
    "rt_raycast_screen_bounds_n": (C.c_int, [_FP, C.POINTER(C.c_double), C.POINTER(C.c_double), _I32, _I32, _I32, C.POINTER(C.c_int)]),
    "rt_bvh_node_bytes": (_I64, [_I64]),
    "rt_bvh_tri_bytes": (_I64, [_I64]),
    "rt_bvh_scratch_bytes": (_I64, [_I64]),
    "rt_bvh_build": (C.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _I32, _VP]),
    "rt_raycast_rays": (C.c_int, [_VP, _VP, _I64, _VP, _I64, _VP, _VP]),
    "rt_raycast_primary": (C.c_int, [_VP, _VP, _I64, _VP, _VP, _VP, _FP, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _U64, _VP, _VP,
                                     _I64, _VP, C.POINTER(C.c_int), _I32, _VP, C.POINTER(C.c_int), _VP]),
    "rt_raycast_view_node_bytes": (_I64, [_I64]),
    "rt_raycast_set_view_refit": (C.c_int, [_I32]),
    "rt_camera_frame": (C.c_int, [_FP, _FP, _FP, _FP]),
    "rt_raycast_screen_bounds": (C.c_int, [_FP, C.POINTER(C.c_double), C.POINTER(C.c_double), _I32, _I32, C.POINTER(C.c_int)]),
    "rt_dsl_compile": (C.c_int, [C.c_char_p, C.POINTER(_U64), C.c_char_p, _I32]),
    "rt_dsl_launch": (C.c_int, [_U64, C.c_char_p, _I64, C.POINTER(_VP), _VP]),
    "rt_dsl_unload": (C.c_int, [_U64]),
    "rt_obj_load": (C.c_int, [C.c_char_p, C.POINTER(_U64)]),
    "rt_obj_mesh_count": (C.c_int, [_U64]),
    "rt_obj_mesh_info": (C.c_int, [_U64, _I32, C.POINTER(_I64), C.POINTER(_I64), C.POINTER(C.c_int)]),
    "rt_obj_mesh_rows": (C.c_int, [_U64, _I32, _FP, _I64]),
    "rt_obj_free": (C.c_int, [_U64]),
    "rt_peer_alloc": (C.c_int, [_I64, C.POINTER(_VP)]),
    "rt_peer_free": (C.c_int, [_VP]),
    "rt_peer_export": (C.c_int, [_VP, _VP]),
    "rt_peer_open": (C.c_int, [_VP, C.POINTER(_VP)]),
    "rt_peer_close": (C.c_int, [_VP]),
    "rt_copy_rect": (C.c_int, [_VP, _I64, _VP, _I64, _I64, _I64, _VP]),
    "rt_push_tiles_state_bytes": (_I64, [_I32, _I32]),
    "rt_push_tiles": (C.c_int, [_VP, _VP, _I32, _I32, _U32, _VP, _VP, _VP]),
    "rt_copy_stripes": (C.c_int, [_VP, _VP, _I64, _I64, _I64, _I64, _I64, _I32, _I32, _I32, _VP]),
    "rt_raster_points_scratch_bytes": (_I64, [_I64]),
    "rt_raster_draw_points": (C.c_int, [_VP, _VP, _VP, _I64, _I32, _FP, _U64, _I32, _I32, _VP, _VP, _I64, _VP, _FP, _I32, _U32, C.POINTER(C.c_int), _VP]),
    "rt_texture_create": (C.c_int, [_VP, _I32, _I32, C.POINTER(_U64)]),
    "rt_texture_destroy": (C.c_int, [_U64]),
}


class NativeUnavailable(RuntimeError):
    pass


class NativeError(RuntimeError):
    pass


_lib = None


def lib():
    """Load the shared library once; raise loudly if it is not there."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeUnavailable(
                f"{LIB_PATH} is not built. Run `python -m rendertoy_b200.build` (needs nvcc). "
                "rendertoy_b200 has no CPU fallback.")
        try:
            handle = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise NativeUnavailable(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        if handle.rt_abi_version() != 1:
            raise NativeUnavailable("librendertoy_b200.so ABI version mismatch; rebuild")
        _lib = handle
    return _lib


def check(rc):
    if rc != 0:
        raise NativeError(f"librendertoy_b200 error {rc}: {lib().rt_last_error().decode(errors='replace')}")


def call(name, *args):
    check(getattr(lib(), name)(*args))


def device_info():
    v = [C.c_int() for _ in range(4)]
    call("rt_device_info", *[C.byref(x) for x in v])
    return {"sm_count": v[0].value, "l2_bytes": v[1].value, "cc": (v[2].value, v[3].value)}


def float_array(values):
    return (C.c_float * len(values))(*[float(x) for x in values])


_F4_CACHE = {}


def float4_const(v):
    """Cached (c_float * 4)(v, v, v, v)."""
    a = _F4_CACHE.get(v)
    if a is None:
        a = _F4_CACHE[v] = (C.c_float * 4)(v, v, v, v)
    return a


def float_array_from_bytes(raw, n):
    """(c_float * n) filled from a numpy uint8 view, one memmove."""
    a = (C.c_float * n)()
    C.memmove(a, raw.ctypes.data, 4 * n)
    return a
