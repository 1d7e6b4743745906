"""tools/cuda_emu/check_raster.py -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.

    python tools/cuda_emu/check_raster.py [n_tris] [width height]

Runs the source of csrc/rt_raster.cu (raster_kernel, coverage_kernel, resolve_kernel, the clears, draw_points) on CPU threads
(tools/cuda_emu) and compares depth words and BGRA8 bytes with the oracle's restatement of the reference pipeline, bit for
bit: lesson08 and lesson09 (texture), a camera whose triangles cross the near plane (clipping, second output triangles,
large primitives through the work queue), two composing draws, points.  Kernel logic only; the B200 parity suite is tests/*_gpu.py.
"""
import ctypes as C
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
os.environ.setdefault("RENDERTOY_B200_HOST_BUFFERS", "1")


def main(n_tris=1200, W=96, H=64, quick=False):
    import emu_build
    import oracle
    from oracle import host_math as hm
    from rendertoy_b200 import scenes
    oracle.build()
    L = C.CDLL(emu_build.build("rt_raster"))
    VP, I64, I32, U32, U64, FP = C.c_void_p, C.c_int64, C.c_int, C.c_uint32, C.c_uint64, C.POINTER(C.c_float)
    L.rt_raster_scratch_bytes.restype = I64
    L.rt_raster_scratch_bytes.argtypes = [I32, I64, I32, I32]
    L.rt_raster_points_scratch_bytes.restype = I64
    L.rt_raster_points_scratch_bytes.argtypes = [I64]
    sig = [VP, VP, VP, I64, I32, FP, U64, I32, I32, VP, VP, I64, VP, FP, I32, U32, C.POINTER(C.c_int), VP]
    L.rt_raster_draw_triangles.argtypes = sig
    L.rt_raster_draw_points.argtypes = sig
    L.rt_texture_create.argtypes = [VP, I32, I32, C.POINTER(U64)]
    L.rt_last_error.restype = C.c_char_p

    rng = np.random.default_rng(5)
    tex = np.ones((19, 13, 4), np.float32)
    tex[:, :, 0:3] = rng.integers(0, 256, (19, 13, 3)) / 255.0
    texels = np.zeros(19 * 13 * 4 + 256, np.float32)                       # room to find a 512-byte aligned start
    off = (-texels.ctypes.data % 512) // 4
    texels[off:off + tex.size] = tex.ravel()
    handle = U64(0)
    assert L.rt_texture_create(texels.ctypes.data + 4 * off, 13, 19, C.byref(handle)) == 0, L.rt_last_error()

    def soa(rows):
        n = rows.shape[0]
        pos4 = np.ones((n, 4), np.float32); pos4[:, :3] = rows[:, 0:3]
        nrm4 = np.zeros((n, 4), np.float32); nrm4[:, :3] = rows[:, 4:7]
        return pos4, nrm4

    a = scenes.dragon(n_tris)
    b = scenes.dragon(max(n_tris // 2, 60), seed=3)
    b[:, 0:3] = b[:, 0:3] * np.float32(0.8) + np.float32(0.05)
    cams = {"lesson camera": (hm.rotate(0.5, (0, 1, 0)), hm.look_at((0, 0.3, 1.0), (0, 0, 0), (0, 1, 0))),
            "far, odd phase": (hm.rotate(2.1, (0, 1, 0)), hm.look_at((0, 0.3, 2), (0, 0, 0), (0, 1, 0))),
            "inside the mesh (near-plane clipping)": (hm.rotate(4.5, (0, 1, 0)), hm.look_at((0.12, 0.32, 0.3), (0, 0, 0), (0, 1, 0)))}
    ok = True
    if quick:       # the CPU test suite: the lesson camera and the clipping one
        cams.pop("far, odd phase")
    for cname, (Wm, Vm) in cams.items():
        P = hm.perspective(aspect_ratio=W / H)
        gl = np.concatenate([Wm.ravel(), Vm.ravel(), P.ravel()]).astype(np.float32)
        for shader in (8, 9):
            for mode in ("two draws", "points"):
                if quick and mode == "points" and (shader == 9 or cname != "lesson camera"):
                    continue
                t0 = time.time()
                key = np.zeros(W * H, np.uint64)
                bgra = np.full((H, W), 0x55555555, np.uint32)
                clear = (C.c_float * 4)(0, 0, 0, 0)
                od = ob = None
                first = True
                for rows in (a, b):
                    pos4, nrm4 = soa(rows)
                    n = rows.shape[0] // 3 if mode != "points" else rows.shape[0]
                    if mode == "points":
                        nb = int(L.rt_raster_points_scratch_bytes(n)); fn = L.rt_raster_draw_points; oracle_draw = oracle.draw_points
                    else:
                        nb = int(L.rt_raster_scratch_bytes(shader, n, W, H)); fn = L.rt_raster_draw_triangles; oracle_draw = oracle.draw_triangles
                    scratch = np.zeros(nb, np.uint8)
                    rc = fn(pos4.ctypes.data, nrm4.ctypes.data, None, n, shader, gl.ctypes.data_as(FP), handle.value if shader == 9 else 0, W, H,
                            key.ctypes.data, scratch.ctypes.data, nb, bgra.ctypes.data, clear if first else None, 1 if first else 0,
                            0x3F800000, None, None)
                    assert rc == 0, L.rt_last_error()
                    r = oracle_draw(shader, W, H, rows, gl, texture=tex if shader == 9 else None, depth=od, bgra=ob)
                    od, ob = r.depth, r.bgra
                    first = False
                depth = (key >> np.uint64(32)).astype(np.uint32).reshape(H, W)
                same_d = np.array_equal(depth, od)
                same_c = np.array_equal(bgra.view(np.uint8).reshape(H, W, 4), ob)
                ok &= same_d and same_c
                print(f"{cname:40s} lesson{shader:02d} {mode:9s} depth {'==' if same_d else '!='} oracle, colour {'==' if same_c else '!='} "
                      f"({int((od != 0x3F800000).sum())} of {W * H} pixels covered, {time.time() - t0:.1f} s)", flush=True)
    # ---- image-space partition: every rank draws the whole mesh with its `owner`; owned pixels must equal the full-frame run,
    # everything else must stay untouched (sentinels).  Stripes of 8 rows over 3 ranks inside a rect that cuts the mesh, on the
    # clipping camera (large primitives -> work queue, second output triangles) and the lesson camera.
    def draw_all(gl, shader, owner, key, bgra, queue_items=None):
        first = True
        for rows in (a, b):
            pos4, nrm4 = soa(rows)
            n = rows.shape[0] // 3
            nb = int(L.rt_raster_scratch_bytes(shader, n, W, H))
            if queue_items is not None:     # a work queue of only `queue_items` items: most reservations fail and fall back inline
                nb = 256 + 2 * n * (4 if shader == 8 else 6) * 16 + 104 * queue_items
            scratch = np.zeros(nb, np.uint8)
            rc = L.rt_raster_draw_triangles(pos4.ctypes.data, nrm4.ctypes.data, None, n, shader, gl.ctypes.data_as(FP), handle.value if shader == 9 else 0,
                                            W, H, key.ctypes.data, scratch.ctypes.data, nb, bgra.ctypes.data, (C.c_float * 4)(0, 0, 0, 0) if first else None,
                                            1 if first else 0, 0x3F800000, owner, None)
            assert rc == 0, L.rt_last_error()
            first = False

    for cname, (Wm, Vm) in cams.items():
        P = hm.perspective(aspect_ratio=W / H)
        gl = np.concatenate([Wm.ravel(), Vm.ravel(), P.ravel()]).astype(np.float32)
        for shader in ((8,) if quick else (8, 9)):
            t0 = time.time()
            key_full = np.zeros(W * H, np.uint64); bgra_full = np.full((H, W), 0x55555555, np.uint32)
            draw_all(gl, shader, None, key_full, bgra_full)
            if shader == 8:      # overflowing large-primitive queue (5 and 40 items): same frame, no holes, no stale items
                for cap in (5, 40):
                    key_q = np.zeros(W * H, np.uint64); bgra_q = np.full((H, W), 0x55555555, np.uint32)
                    draw_all(gl, shader, None, key_q, bgra_q, queue_items=cap)
                    same = np.array_equal(key_q, key_full) and np.array_equal(bgra_q, bgra_full)
                    ok &= same
                    print(f"{cname:40s} lesson{shader:02d} work queue of {cap} items: frame {'==' if same else '!='} the unconstrained draw", flush=True)
            rect = (W // 7, H // 9, W - W // 5, H - H // 6)
            mod = 3
            key_sum = np.full(W * H, 0x1234567812345678, np.uint64); bgra_sum = np.full((H, W), 0xABABABAB, np.uint32)
            yy, xx = np.mgrid[0:H, 0:W]
            in_rect = (xx >= rect[0]) & (xx <= rect[2]) & (yy >= rect[1]) & (yy <= rect[3])
            good = True
            for rem in range(mod):
                key = np.full(W * H, 0x1234567812345678, np.uint64); bgra = np.full((H, W), 0xABABABAB, np.uint32)
                draw_all(gl, shader, (C.c_int * 7)(*rect, 8, mod, rem), key, bgra)
                owned = in_rect & ((yy // 8) % mod == rem)
                k2 = key.reshape(H, W)
                good &= bool(np.all(k2[~owned] == np.uint64(0x1234567812345678))) and bool(np.all(bgra[~owned] == 0xABABABAB))
                key_sum.reshape(H, W)[owned] = k2[owned]; bgra_sum[owned] = bgra[owned]
            good &= np.array_equal(key_sum.reshape(H, W)[in_rect], key_full.reshape(H, W)[in_rect]) and np.array_equal(bgra_sum[in_rect], bgra_full[in_rect])
            ok &= good
            print(f"{cname:40s} lesson{shader:02d} 3-rank stripes in a rect: owned pixels {'==' if good else '!='} the full-frame draw, the rest untouched "
                  f"({time.time() - t0:.1f} s)", flush=True)
    print("RASTER BIT-EXACT" if ok else "RASTER MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    a = [int(x) for x in sys.argv[1:]]
    sys.exit(main(*(a[:1] or [1200]), *(a[1:3] if len(a) >= 3 else (96, 64))))
