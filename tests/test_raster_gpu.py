"""GPU parity: rendering.Raster (native sm_100a kernels through the C ABI) vs the oracle's restatement of the
reference pipeline, on the same seeded inputs.  Depth bits and BGRA8 bytes must be identical."""
import numpy as np
import pytest

from rendertoy_b200 import lessons, scenes

pytestmark = pytest.mark.gpu


def _compare(raster, result, label):
    depth = raster.get_depth_buffer().get().reshape(result.depth.shape)
    bgra = raster.get_render_target().get()
    n = depth.size
    bad_depth = int((depth != result.depth).sum())
    bad_rgb = int((bgra != result.bgra).any(axis=-1).sum())
    covered = int((result.winner != 0xFFFFFFFF).sum())
    print(f"{label}: pixels={n} covered={covered} depth_mismatch={bad_depth} colour_mismatch={bad_rgb} stats={result.stats}")
    assert bad_depth == 0, f"{label}: {bad_depth} depth words differ"
    assert bad_rgb == 0, f"{label}: {bad_rgb} BGRA8 pixels differ"
    assert covered > 0


def _upload(ren, rows):
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    return vb


def _texture(seed=3, w=50, h=37):
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)


@pytest.mark.parametrize("lesson,width,height,n_tris,t", [
    (8, 640, 480, 20_000, 0.5),
    (8, 1920, 1080, 100_000, 0.5),
    (8, 333, 211, 5_000, 2.1),
    (9, 640, 480, 20_000, 0.5),
    (9, 1920, 1080, 100_000, 1.3),
])
def test_lesson_scene_matches_oracle(ren, oracle, lesson, width, height, n_tris, t):
    rows = scenes.dragon(n_tris)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    presenter = ren.create_presenter(width, height)
    tex = None
    if lesson == 8:
        raster, g = lessons.build_lesson08(ren, presenter.get_render_target())
    else:
        tex = _texture()
        raster, g, _, _ = lessons.build_lesson09(ren, presenter.get_render_target(), tex)
    world, view, proj = scenes.lesson_camera(ren, 8, t, width, height)
    lessons.set_transforms(ren, g, world, view, proj)
    lessons.render_frame(ren, raster, vb)
    texf = None
    if tex is not None:
        texf = np.ones((tex.shape[0], tex.shape[1], 4), np.float32)
        texf[:, :, 0:3] = tex / 255.0
    res = oracle.draw_triangles(lesson, width, height, rows, lessons.globals_as_floats(g), texture=texf)
    _compare(raster, res, f"lesson{lesson:02d} {width}x{height} T={rows.shape[0] // 3}")


def test_deferred_clears_are_not_observable(ren, oracle):
    """clear() of the render target / depth buffer is deferred into the next draw; any other access executes it first."""
    w, h = 96, 64
    pres = ren.create_presenter(w, h)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    target = raster.get_render_target()
    ren.clear(target, np.array([1.0, 0.5, 0.25, 1.0], dtype=np.float32))
    ren.clear(raster.get_depth_buffer(), 0.75)
    assert (target.get() == np.array([64, 128, 255, 255], dtype=np.uint8)).all()            # B, G, R, A; 0.5*255 rounds to even
    assert (raster.get_depth_buffer().get() == np.float32(0.75).view(np.uint32)).all()
    # clear + draw: pixels nobody wins must carry the clear colour, winners the shaded colour
    rows = scenes.dragon(1500)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8, 0.5, w, h))
    ren.clear(target, np.array([0.0, 0.0, 1.0, 1.0], dtype=np.float32))
    ren.clear(raster.get_depth_buffer(), 1.0)
    raster.draw_triangles(vb, None)
    bg = np.zeros((h, w, 4), np.uint8); bg[:, :, 0] = 255; bg[:, :, 3] = 255
    res = oracle.draw_triangles(8, w, h, rows, lessons.globals_as_floats(g), bgra=bg)
    assert np.array_equal(target.get(), res.bgra) and (res.winner == 0xFFFFFFFF).any()
    assert np.array_equal(raster.get_depth_buffer().get().reshape(h, w), res.depth)
    # a second draw without clears composes on top (nothing pending any more)
    raster.draw_triangles(vb, None)
    assert np.array_equal(target.get(), res.bgra)


@pytest.mark.parametrize("lesson,indexed", [(8, False), (9, False), (8, True)])
def test_draw_points_matches_oracle(ren, oracle, lesson, indexed):
    w, h = 320, 240
    rows = scenes.dragon(3000)
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    pres = ren.create_presenter(w, h)
    tex = texf = None
    if lesson == 8:
        raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    else:
        tex = _texture()
        raster, g, _, _ = lessons.build_lesson09(ren, pres.get_render_target(), tex)
        texf = np.ones((tex.shape[0], tex.shape[1], 4), np.float32)
        texf[:, :, 0:3] = tex / 255.0
    idx = np.random.default_rng(4).integers(0, rows.shape[0], 5000).astype(np.int32) if indexed else None
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8, 0.5, w, h))
    ren.clear(raster.get_render_target())
    ren.clear(raster.get_depth_buffer(), 1.0)
    raster.draw_points(vb, None if idx is None else ren.create_buffer_from(idx))
    res = oracle.draw_points(lesson, w, h, rows, lessons.globals_as_floats(g), indices=idx, texture=texf)
    _compare(raster, res, f"points lesson{lesson:02d} indexed={indexed}")
    # points and triangles compose on the same targets
    raster.draw_triangles(vb, None)
    res2 = oracle.draw_triangles(lesson, w, h, rows, lessons.globals_as_floats(g), texture=texf, depth=res.depth, bgra=res.bgra)
    _compare(raster, res2, "points then triangles")


def test_content_rect_holds_everything_drawn(ren):
    """Raster.content_rect: outside it the render target is the clear colour (the contract of the sparse read-back)."""
    from oracle import host_math as hm
    w, h = 640, 360
    rows = scenes.dragon(6000)
    vb = _upload(ren, rows)
    small = rows.copy()
    small[:, 0:3] = small[:, 0:3] * np.float32(0.3) + np.float32([0.6, 0.2, 0.0])
    vb2 = _upload(ren, small)
    pres = ren.create_presenter(w, h)
    raster, g = lessons.build_lesson08(ren, pres.get_render_target())
    assert raster.content_rect == (0, 0, w - 1, h - 1)                      # nothing known yet

    def check(rect):
        img = raster.get_render_target().get().view(np.uint32).reshape(h, w)
        assert img.any()
        outside = np.ones((h, w), bool)
        if rect[2] >= rect[0]:
            outside[rect[1]:rect[3] + 1, rect[0]:rect[2] + 1] = False
        assert not img[outside].any(), "pixels drawn outside content_rect"

    for lesson, t in [(6, 0.4), (8, 1.9), (6, 3.3)]:
        lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, lesson, t, w, h))
        raster.get_render_target().buffer.tensor().fill_(0x77)              # stale pixels everywhere
        lessons.render_frame(ren, raster, vb)
        r1 = raster.content_rect
        assert r1 != (0, 0, w - 1, h - 1) and r1[2] > r1[0]
        check(r1)
        raster.draw_triangles(vb2, None)                                      # second draw of the frame: the union
        r2 = raster.content_rect
        assert r2[0] <= r1[0] and r2[1] <= r1[1] and r2[2] >= r1[2] and r2[3] >= r1[3]
        check(r2)
        raster.draw_points(vb2)
        check(raster.content_rect)
    # the camera inside the mesh: triangles are clipped, no bound
    W, V, P = hm.rotate(4.5, (0, 1, 0)), hm.look_at((0.12, 0.32, 0.3), (0, 0, 0), (0, 1, 0)), hm.perspective(aspect_ratio=w / h)
    lessons.set_transforms(ren, g, *(ren.make_float4x4(np.ascontiguousarray(x)) for x in (W, V, P)))
    lessons.render_frame(ren, raster, vb)
    assert raster.content_rect == (0, 0, w - 1, h - 1)
    # a vertex buffer written after it was drawn: unknown again
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8, 0.5, w, h))
    lessons.render_frame(ren, raster, vb)
    assert raster.content_rect != (0, 0, w - 1, h - 1)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[0, 0] += 0.0
    assert raster.content_rect == (0, 0, w - 1, h - 1)


@pytest.mark.parametrize("lesson,width,height,n_tris,world,cam_lesson", [
    (8, 1920, 1080, 100_000, 8, 8),      # configs[1] split over 8 ranks, stripes of parallel.BAND rows
    (9, 1920, 1080, 100_000, 2, 8),
    (8, 3840, 2160, 100_000, 4, 8),      # the 4K frame of configs[3], frame-filling camera
    (8, 333, 211, 5_000, 3, 6),
])
def test_stripe_partition_equals_the_whole_frame(ren, lesson, width, height, n_tris, world, cam_lesson):
    """SURVEY.md 8e "bin only into owned tiles": every rank draws the whole mesh with set_scissor(stripes=(BAND, world, rank)).
    The union of the ranks' owned pixels must equal the unpartitioned draw byte for byte (depth words and BGRA8), and no rank
    may touch a pixel it does not own (sentinels survive).  Two composing draws per frame."""
    from rendertoy_b200 import parallel
    rows_a = scenes.dragon(n_tris)
    rows_b = scenes.dragon(max(n_tris // 4, 100), seed=3)
    rows_b[:, 0:3] = rows_b[:, 0:3] * np.float32(0.8) + np.float32(0.05)
    vbs = [_upload(ren, rows_a), _upload(ren, rows_b)]
    tex = _texture()
    cam = scenes.lesson_camera(ren, cam_lesson, 0.7, width, height)

    def build():
        target = ren.create_presenter(width, height).get_render_target()
        if lesson == 8:
            raster, g = lessons.build_lesson08(ren, target)
        else:
            raster, g, _, _ = lessons.build_lesson09(ren, target, tex)
        lessons.set_transforms(ren, g, *cam)
        return raster

    def frame(raster):
        ren.clear(raster.get_render_target())
        ren.clear(raster.get_depth_buffer(), 1.0)
        for vb in vbs:
            raster.draw_triangles(vb, None)

    whole = build()
    frame(whole)
    depth_ref = whole.get_depth_buffer().get().reshape(height, width)
    bgra_ref = whole.get_render_target().get().view(np.uint32).reshape(height, width)
    assert (depth_ref != 0x3F800000).any()
    depth_sum = np.zeros_like(depth_ref)
    bgra_sum = np.zeros_like(bgra_ref)
    yy = np.arange(height)[:, None]
    for rank in range(world):
        part = build()
        # sentinels: a depth the draw would beat everywhere, a colour no shader produces
        part.get_depth_buffer().set(np.full(width * height, 0x7F7F7F7F, np.uint32))
        with ren.mapped(part.get_render_target()) as m:
            m.view(np.uint32)[...] = 0xABCDEF01
        part.set_scissor(stripes=(parallel.BAND, world, rank))
        frame(part)
        d = part.get_depth_buffer().get().reshape(height, width)
        c = part.get_render_target().get().view(np.uint32).reshape(height, width)
        owned = np.broadcast_to((yy // parallel.BAND) % world == rank, d.shape)
        assert (d[~owned] == 0x7F7F7F7F).all() and (c[~owned] == 0xABCDEF01).all(), f"rank {rank} wrote outside its stripes"
        depth_sum[owned] = d[owned]
        bgra_sum[owned] = c[owned]
    assert np.array_equal(depth_sum, depth_ref), f"{int((depth_sum != depth_ref).sum())} depth words differ from the whole frame"
    assert np.array_equal(bgra_sum, bgra_ref), f"{int((bgra_sum != bgra_ref).sum())} pixels differ from the whole frame"


def test_scissor_rect_and_points(ren, oracle):
    """A plain scissor rect that cuts the mesh: inside == the oracle's full frame, outside untouched; draw_points too;
    content_rect is clipped to it; lifting the scissor restores whole-frame draws."""
    w, h = 640, 480
    rows = scenes.dragon(20_000)
    vb = _upload(ren, rows)
    raster, g = lessons.build_lesson08(ren, ren.create_presenter(w, h).get_render_target())
    lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8, 0.5, w, h))
    rect = (201, 97, 433, 350)
    yy, xx = np.mgrid[0:h, 0:w]
    inside = (xx >= rect[0]) & (xx <= rect[2]) & (yy >= rect[1]) & (yy <= rect[3])
    for points in (False, True):
        draw = raster.draw_points if points else raster.draw_triangles
        ref = (oracle.draw_points if points else oracle.draw_triangles)(8, w, h, rows, lessons.globals_as_floats(g))
        raster.set_scissor(rect=rect)
        raster.get_depth_buffer().set(np.full(w * h, 0x7F7F7F7F, np.uint32))
        with ren.mapped(raster.get_render_target()) as m:
            m.view(np.uint32)[...] = 0xABCDEF01
        ren.clear(raster.get_render_target())
        ren.clear(raster.get_depth_buffer(), 1.0)
        draw(vb, None)
        d = raster.get_depth_buffer().get().reshape(h, w)
        c = raster.get_render_target().get()
        assert np.array_equal(d[inside], ref.depth[inside]) and np.array_equal(c[inside], ref.bgra[inside])
        assert (d[~inside] == 0x7F7F7F7F).all() and (c.view(np.uint32).reshape(h, w)[~inside] == 0xABCDEF01).all()
        cr = raster.content_rect
        assert cr[0] >= rect[0] and cr[1] >= rect[1] and cr[2] <= rect[2] and cr[3] <= rect[3]
        raster.set_scissor()
        ren.clear(raster.get_render_target())
        ren.clear(raster.get_depth_buffer(), 1.0)
        draw(vb, None)
        assert np.array_equal(raster.get_depth_buffer().get().reshape(h, w), ref.depth)
        assert np.array_equal(raster.get_render_target().get(), ref.bgra)


def test_draw_frame_equals_the_four_tutorial_calls(ren, oracle):
    """Raster.draw_frame(vb, ib, transforms) == mapped(globals) + clear + clear + draw_triangles, bit for bit, and leaves the
    globals buffer holding the same matrices; with transforms=None it draws with what the buffer holds."""
    import ctypes
    w, h = 640, 360
    rows = scenes.dragon(8_000)
    vb = _upload(ren, rows)
    cam = scenes.lesson_camera(ren, 8, 1.1, w, h)
    a, ga = lessons.build_lesson08(ren, ren.create_presenter(w, h).get_render_target())
    lessons.set_transforms(ren, ga, *cam)
    lessons.render_frame(ren, a, vb)
    ref = oracle.draw_triangles(8, w, h, rows, lessons.globals_as_floats(ga))
    g48 = lessons.globals_as_floats(ga)
    for transforms in (cam, g48, (ctypes.c_float * 48)(*g48.tolist())):
        b, gb = lessons.build_lesson08(ren, ren.create_presenter(w, h).get_render_target())
        b.get_depth_buffer().set(np.full(w * h, 0x12345678, np.uint32))       # stale contents the frame's clears must remove
        b.draw_frame(vb, None, transforms)
        assert np.array_equal(lessons.globals_as_floats(gb), g48)
        assert np.array_equal(b.get_depth_buffer().get().reshape(h, w), ref.depth)
        assert np.array_equal(b.get_render_target().get(), ref.bgra)
        b.draw_frame(vb, None)                                                  # same matrices, from the buffer
        assert np.array_equal(b.get_render_target().get(), ref.bgra)


def test_frame_store_tile_push(ren):
    """FrameStore.push_tiles (rt_push_tiles): an orbit of raster frames cycling through two slots of a cleared store; after
    every push the slot equals the frame bit for bit although only the non-clear 32x32 tiles (and the ones that were non-clear
    in the slot before) were stored; a frame size that is not a multiple of the tile exercises the partial tiles."""
    import torch
    from rendertoy_b200 import parallel
    for w, h, lo_frac in ((1920, 1080, 0.75), (500, 333, 0.95)):
        rows = scenes.dragon(20_000)
        vb = _upload(ren, rows)
        raster, g = lessons.build_lesson08(ren, ren.create_presenter(w, h).get_render_target())
        store = parallel.FrameStore(2, w, h)
        assert store.ok
        try:
            side = torch.cuda.Stream()
            for k in range(10):
                lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8 if k % 3 else 6, 0.6 * k, w, h))
                lessons.render_frame(ren, raster, vb)
                side.wait_stream(torch.cuda.current_stream())
                store.push_tiles(k % 2, raster.get_render_target().ptr, side.cuda_stream)
                torch.cuda.current_stream().wait_stream(side)
                frame = raster.get_render_target().buffer.tensor().view(torch.int32).view(h, w)
                assert torch.equal(store.frames()[k % 2], frame), f"slot differs from frame {k}"
                assert frame.any()
            moved = int(store.tile_bytes.item())
            assert 0 < moved < lo_frac * 10 * ((w + 31) // 32) * ((h + 31) // 32) * 4096, moved
        finally:
            store.close()


def test_tile_readback_into_pinned_host_memory(ren):
    """parallel.TileFrameCopier with a PINNED HOST destination (the e2e read-back): the rt_push_tiles kernel stores the non-clear
    tiles of each frame straight into host memory; after every frame the host frame equals the device frame, with fewer bytes
    crossing PCIe than the frames hold."""
    import torch
    from rendertoy_b200 import parallel
    from rendering._core import stream_ptr
    w, h = 1920, 1080
    rows = scenes.dragon(20_000)
    vb = _upload(ren, rows)
    raster, g = lessons.build_lesson08(ren, ren.create_presenter(w, h).get_render_target())
    host = [torch.zeros((h, w), dtype=torch.int32).pin_memory() for _ in range(2)]
    copier = parallel.TileFrameCopier(w, h)
    for k in range(8):
        lessons.set_transforms(ren, g, *scenes.lesson_camera(ren, 8 if k % 3 else 6, 0.6 * k, w, h))
        lessons.render_frame(ren, raster, vb)
        copier.copy(k % 2, host[k % 2].data_ptr(), raster.get_render_target().ptr, stream_ptr())
        torch.cuda.synchronize()
        frame = raster.get_render_target().buffer.tensor().view(torch.int32).view(h, w).cpu()
        assert torch.equal(host[k % 2], frame), f"host frame differs from frame {k}"
        assert frame.any()
    assert 0 < copier.bytes_moved() < 0.7 * 8 * w * h * 4
