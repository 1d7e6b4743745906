"""tools/cuda_emu/check_peer.py -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Runs the source of csrc/rt_peer.cu on the host (tools/cuda_emu: "device" memory is host memory, copies are synchronous):
  rt_copy_stripes  (the tile partition's gather: one 3-D copy + up to two 2-D ones) against a row-by-row numpy definition, over
                   random frames, rects, stripe heights and rank counts, including stripes cut by the rect and by the frame end;
  rt_push_tiles    (the tile-sparse gather kernel) over a cycle of random frames through two destinations: afterwards the
                   destination equals the source, only tiles that are (or were) non-clear are stored, the byte counter adds up;
  rt_copy_rect, rt_peer_alloc/export/open (plumbing).
"""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def main(quick=False):
    import emu_build
    L = C.CDLL(emu_build.build("rt_peer"))
    VP, I64, I32, U32 = C.c_void_p, C.c_int64, C.c_int, C.c_uint32
    L.rt_copy_stripes.argtypes = [VP, VP, I64, I64, I64, I64, I64, I32, I32, I32, VP]
    L.rt_copy_rect.argtypes = [VP, I64, VP, I64, I64, I64, VP]
    L.rt_push_tiles.argtypes = [VP, VP, I32, I32, U32, VP, VP, VP]
    L.rt_push_tiles_state_bytes.restype = I64
    L.rt_push_tiles_state_bytes.argtypes = [I32, I32]
    L.rt_peer_alloc.argtypes = [I64, C.POINTER(VP)]
    L.rt_peer_export.argtypes = [VP, VP]
    L.rt_peer_open.argtypes = [VP, C.POINTER(VP)]
    L.rt_last_error.restype = C.c_char_p
    rng = np.random.default_rng(23)
    ok = True

    # ---- rt_copy_stripes
    n_cases = 150 if quick else 600
    for case in range(n_cases):
        W, H = int(rng.integers(8, 90)), int(rng.integers(5, 140))
        rows, mod = int(rng.integers(1, 20)), int(rng.integers(1, 6))
        rem = int(rng.integers(0, mod))
        x0, x1 = sorted(int(v) for v in rng.integers(0, W, 2))
        y0, y1 = sorted(int(v) for v in rng.integers(0, H, 2))
        if case % 7 == 0:
            y0, y1 = 0, H - 1
        src = rng.integers(1, 2 ** 32, (H, W), dtype=np.uint32)
        dst = np.zeros((H, W), np.uint32)
        want = dst.copy()
        for y in range(y0, y1 + 1):
            if (y // rows) % mod == rem:
                want[y, x0:x1 + 1] = src[y, x0:x1 + 1]
        rc = L.rt_copy_stripes(dst.ctypes.data, src.ctypes.data, 4 * W, 4 * x0, 4 * (x1 - x0 + 1), y0, y1, rows, mod, rem, None)
        assert rc == 0, L.rt_last_error()
        if not np.array_equal(dst, want):
            ok = False
            print(f"rt_copy_stripes MISMATCH: {W}x{H} rect ({x0},{y0})-({x1},{y1}) stripes ({rows},{mod},{rem})")
    print(f"rt_copy_stripes: {n_cases} random cases {'==' if ok else '!='} the row-by-row definition", flush=True)

    # ---- rt_push_tiles
    for (W, H, clear) in ((96, 64, 0), (100, 70, 0), (64, 40, 0xFF000000)):
        nt = int(L.rt_push_tiles_state_bytes(W, H))
        assert nt == ((W + 31) // 32) * ((H + 31) // 32)
        dests = [np.full((H, W), clear, np.uint32) for _ in range(2)]
        states = [np.zeros(nt, np.uint8) if clear == 0 else np.ones(nt, np.uint8) for _ in range(2)]
        counter = np.zeros(1, np.uint64)
        stored_tiles = 0
        good = True
        for it in range(6 if quick else 14):
            src = np.full((H, W), clear, np.uint32)
            if it % 5 != 4:                                   # every fifth frame is empty: the slot must be cleared again
                bx0, bx1 = sorted(int(v) for v in rng.integers(0, W, 2)); by0, by1 = sorted(int(v) for v in rng.integers(0, H, 2))
                blob = rng.integers(1, 2 ** 32, (by1 - by0 + 1, bx1 - bx0 + 1), dtype=np.uint32)
                blob[blob == clear] = 1
                src[by0:by1 + 1, bx0:bx1 + 1] = blob
            d, st = dests[it % 2], states[it % 2]
            before, prev = d.copy(), st.copy()
            rc = L.rt_push_tiles(d.ctypes.data, src.ctypes.data, W, H, clear, st.ctypes.data, counter.ctypes.data, None)
            assert rc == 0, L.rt_last_error()
            good &= np.array_equal(d, src)
            # tile bookkeeping: the new state says which tiles hold something; a tile that is clear now and was clear before is not stored
            tiles_x = (W + 31) // 32
            for t in range(nt):
                ty, tx = divmod(t, tiles_x)
                blk = src[ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32]
                any_now = bool((blk != clear).any())
                if any_now or prev[t]:
                    stored_tiles += 1
                    good &= int(st[t]) == int(any_now)
                else:
                    good &= int(st[t]) == int(prev[t]) and np.array_equal(d[ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32], before[ty * 32:(ty + 1) * 32, tx * 32:(tx + 1) * 32])
        good &= int(counter[0]) == 4096 * stored_tiles
        ok &= good
        print(f"rt_push_tiles {W}x{H} clear {clear:#x}: destination {'==' if good else '!='} source after every push, {stored_tiles} tiles stored "
              f"({int(counter[0])} bytes counted)", flush=True)

    # ---- plumbing: alloc (zero-filled) / export / open / copy_rect
    p, q = VP(), VP()
    assert L.rt_peer_alloc(4096, C.byref(p)) == 0
    h = C.create_string_buffer(64)
    assert L.rt_peer_export(p, h) == 0 and L.rt_peer_open(h, C.byref(q)) == 0 and p.value == q.value
    mem = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), (4096,))
    ok &= not mem.any()
    a = rng.integers(0, 255, (16, 64), dtype=np.uint8); b = np.zeros_like(a)
    assert L.rt_copy_rect(b.ctypes.data + 8, 64, a.ctypes.data + 8, 64, 24, 10, None) == 0
    ok &= np.array_equal(b[:10, 8:32], a[:10, 8:32]) and not b[10:].any() and not b[:, :8].any() and not b[:, 32:].any()
    print("PEER PLUMBING EXACT" if ok else "PEER MISMATCH")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main(quick="quick" in sys.argv))
