"""Generic kernel DSL (kernel_main / kernel_function / kernel_struct on NVRTC): the tutorials' lesson01-07 kernels.
CPU: the OpenCL C text compiles for sm_100a (NVRTC cross-compiles without a GPU).  GPU: results vs numpy and vs the
golden run of the reference's own lesson06 kernel (tests/golden/dsl_lesson06.npz, oracle/clshim/make_golden.py)."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "dsl_lesson06.npz")


@pytest.fixture(scope="module")
def kernels(ren):
    """Kernel bodies as written in tutorials/lesson01, 02, 04, 06, 07 (comments dropped)."""
    k = {}

    @ren.kernel_main
    def compute(x: [np.float32], y: [np.float32]):
        """
        y[thread_id] = sin(x[thread_id]);
        """
    k["compute"] = compute

    @ren.kernel_main
    def transform(x: [ren.float3], T: ren.float4x4, y: [ren.float3]):
        """
        float4 p = (float4)(x[thread_id], 1.0f);
        p = mul(p, T);
        y[thread_id] = p.xyz / p.w;
        """
    k["transform"] = transform

    @ren.kernel_struct
    class MandelbrotInfo:
        C: ren.float2
        N: np.int32
    k["MandelbrotInfo"] = MandelbrotInfo

    @ren.kernel_function
    def get_color(m: np.float32) -> ren.float4:
        """
        m = min(m, 10.0f);
        float s = 2*(1.0f / (1 + exp(-m)) - 0.5f);
        return (float4)(0.0f, 1.0f-s, fmod(s+0.5,1.0), 1.0f);
        """

    @ren.kernel_main
    def compute_mandelbrot(im: ren.w_image2d_t, info: MandelbrotInfo):
        """
        int2 dim = get_image_dim(im);
        int px = thread_id % dim.x;
        int py = thread_id / dim.x;
        float2 Z = ((float2)((px + 0.5f)/dim.x, (py + 0.5f)/dim.y)) * 2.0f - 1.0f;
        for (int i=0; i<info.N; i++)
            Z = (float2)(Z.x*Z.x - Z.y*Z.y, 2*Z.x*Z.y) + info.C;
        float m = sqrt(dot(Z, Z));
        write_imagef(im, (int2)(px,py), get_color(m));
        """
    k["compute_mandelbrot"] = compute_mandelbrot

    @ren.kernel_struct
    class SplatTransforms:
        World: ren.float4x4
        View: ren.float4x4
        Proj: ren.float4x4
    k["SplatTransforms"] = SplatTransforms

    @ren.kernel_main
    def splat(im: ren.w_image2d_t, vertices: [ren.MeshVertex], info: SplatTransforms):
        """
        int2 dim = get_image_dim(im);
        float3 P = vertices[thread_id].P;
        float3 C = vertices[thread_id].N * 0.5f + 0.5f;
        float4 H = (float4)(P.x, P.y, P.z, 1.0);
        H = mul(H, info.World);
        H = mul(H, info.View);
        H = mul(H, info.Proj);
        H.xyz /= H.w;
        if (any(H.xyz < (float3)(-1.0, -1.0, 0.0)) || any(H.xyz >= 1))
        return;
        int px = (int)(dim.x * (H.x * 0.5 + 0.5));
        int py = (int)(dim.y * (0.5 - H.y * 0.5));
        write_imagef(im, (int2)(px,py), (float4)(C.x, C.y, C.z, 1.0));
        """
    k["splat"] = splat

    @ren.kernel_function
    def C_n_k(n: int, k: int) -> int:
        """
        if (k < n - k) k = n - k;
        long f = 1;
        for (int i = k + 1; i <= n; i++)
            f *= i;
        for (int i = 2; i <= n - k; i++)
            f /= i;
        return (int)f;
        """

    @ren.kernel_main
    def perform_parametric_transform(vertices: [ren.MeshVertex], cps: [ren.float3], cp_count: int):
        """
        float2 uv = vertices[thread_id].C;
        float u = uv.x;
        float v = uv.y;
        float t = u;
        float3 p = (float3)(0,0,0);
        int n = cp_count - 1;
        for (int k = 0; k <= n; k++)
            p += cps[k] * C_n_k(n, k) * pow(t, (float)k) * pow(1 - t, (float)(n - k));
        float4x4 rot = rotation(v * 3.141593 * 2, (float3)(0,1,0));
        float4 h = (float4)(p.x, p.y, p.z, 1.0);
        h = mul(h, rot);
        float3 position = h.xyz;
        vertices[thread_id].P = position;
        """
    k["perform_parametric_transform"] = perform_parametric_transform
    return k


def test_tutorial_kernels_compile_for_sm100a(ren, kernels):
    from rendering import _dsl
    for name in ("compute", "transform", "compute_mandelbrot", "splat", "perform_parametric_transform"):
        src = _dsl.program_source(kernels[name])
        assert _dsl.compile_program(src) != 0
    bad = ren._core.build_kernel_main("broken", {"x": [np.float32]}, "x[thread_id] = undefined_symbol;")
    with pytest.raises(RuntimeError, match="failed to build"):
        _dsl.compile_program(_dsl.program_source(bad))


@pytest.mark.gpu
def test_lesson01_02_math(ren, kernels):
    x = np.linspace(0, 6.0, 1000, dtype=np.float32)
    xb, yb = ren.create_buffer_from(x), ren.create_buffer(1000, np.float32)
    kernels["compute"][1000](xb, yb)
    assert np.allclose(yb.get(), np.sin(x), atol=2e-6)
    pts = ren.create_buffer(64, ren.float3)
    rng = np.random.default_rng(0)
    p = rng.uniform(-1, 1, (64, 3)).astype(np.float32)
    with ren.mapped(pts) as m:
        m.view(np.float32).reshape(64, 4)[:, :3] = p
    out = ren.create_buffer(64, ren.float3)
    T = ren.matmul(ren.rotate(0.3, ren.make_float3(0, 1, 0)), ren.translate(1.0, 2.0, 3.0))
    kernels["transform"][64](pts, np.array(T, dtype=ren.float4x4), out)
    M = ren.to_array(np.array(T, dtype=ren.float4x4)).astype(np.float64)
    exp = np.concatenate([p, np.ones((64, 1))], axis=1) @ M
    got = out.get().view(np.float32).reshape(64, 4)[:, :3]
    assert np.allclose(got, exp[:, :3] / exp[:, 3:4], atol=1e-5)


@pytest.mark.gpu
def test_lesson04_mandelbrot_image(ren, kernels):
    w, h = 64, 48
    im = ren.create_image2d(w, h, ren._core.RGBA)
    info = ren.create_struct(kernels["MandelbrotInfo"])
    with ren.mapped(info) as m:
        m["C"] = ren.make_float2(-0.4, 0.6)
        m["N"] = 20
    kernels["compute_mandelbrot"][w * h](im, info)
    got = im.get()
    px, py = np.meshgrid(np.arange(w), np.arange(h))
    z = ((px + 0.5) / w * 2 - 1) + 1j * ((py + 0.5) / h * 2 - 1)
    with np.errstate(over="ignore", invalid="ignore"):
        for _ in range(20):
            z = z * z + (-0.4 + 0.6j)
        mm = np.minimum(np.nan_to_num(np.abs(z), nan=np.inf), 10.0)
    s = 2 * (1.0 / (1 + np.exp(-mm)) - 0.5)
    exp_g = np.clip(np.rint((1.0 - s) * 255), 0, 255)
    assert (got[:, :, 3] == 255).all() and (got[:, :, 2] == 0).all()          # A = 1, R = 0 (bytes are B, G, R, A)
    assert (np.abs(got[:, :, 1].astype(int) - exp_g) <= 1).mean() > 0.98


@pytest.mark.gpu
def test_lesson06_point_splat_matches_reference_run(ren, kernels):
    z = np.load(GOLDEN)
    rows, w, h = z["rows"], int(z["width"]), int(z["height"])
    vb = ren.create_buffer(rows.shape[0], ren.MeshVertex)
    with ren.mapped(vb) as m:
        m.view(np.float32).reshape(rows.shape)[:] = rows
    im = ren.create_image2d(w, h, ren._core.RGBA)
    info = ren.create_struct(kernels["SplatTransforms"])
    gl = z["globals"].reshape(3, 16)
    with ren.mapped(info) as m:
        m["World"], m["View"], m["Proj"] = (ren.make_float4x4(np.ascontiguousarray(x)) for x in gl)
    ren.clear(im)
    kernels["splat"][vb.shape](im, vb, info)
    got, ref = im.get(), z["bgra"]
    assert np.array_equal(got[:, :, 3] != 0, ref[:, :, 3] != 0), "different pixels were written"
    same = (got == ref).all(axis=-1)
    assert same.mean() > 0.995        # pixels hit by several vertices keep whichever thread wrote last (a race in the reference too)


@pytest.mark.gpu
def test_lesson07_parametric_transform(ren, kernels):
    mesh = ren.manifold(8, 6)
    cps = np.array([[0.0, 0.0, 0.0], [0.4, 0.3, 0.0], [0.2, 0.7, 0.0], [0.5, 1.0, 0.0]], np.float32)
    cb = ren.create_buffer(4, ren.float3)
    with ren.mapped(cb) as m:
        m.view(np.float32).reshape(4, 4)[:, :3] = cps
    before = mesh.vertices.get().view(np.float32).reshape(-1, 20).copy()
    kernels["perform_parametric_transform"][mesh.vertices.shape](mesh.vertices, cb, 4)
    after = mesh.vertices.get().view(np.float32).reshape(-1, 20)
    from math import comb
    u, v = before[:, 8].astype(np.float64), before[:, 9].astype(np.float64)
    p = sum(cps[k][None, :].astype(np.float64) * comb(3, k) * (u ** k)[:, None] * ((1 - u) ** (3 - k))[:, None] for k in range(4))
    ang = v * 3.141593 * 2
    c, s = np.cos(ang), np.sin(ang)
    exp = np.stack([p[:, 0] * c + p[:, 2] * s, p[:, 1], -p[:, 0] * s + p[:, 2] * c], axis=1)
    assert np.allclose(after[:, 0:3], exp, atol=2e-5)
    assert np.array_equal(after[:, 4:], before[:, 4:])


@pytest.fixture(scope="module")
def linear_kernel(ren):
    @ren.kernel_struct
    class LinearSampleInfo:
        Tex: ren.Texture2D

    @ren.kernel_main
    def sample_both(nearest: [ren.float4], linear: [ren.float4], coords: [ren.float2], info: LinearSampleInfo):
        """
        nearest[thread_id] = sample2D(info.Tex, coords[thread_id]);
        linear[thread_id] = sample2D_linear(info.Tex, coords[thread_id]);
        """
    return LinearSampleInfo, sample_both


def test_sample2D_linear_compiles(ren, linear_kernel):
    from rendering import _dsl
    ren.create_texture2D(4, 4)      # the program bakes the pool address in: the pool must exist
    assert _dsl.compile_program(_dsl.program_source(linear_kernel[1])) != 0


@pytest.mark.gpu
def test_sample2D_linear_against_numpy(ren, linear_kernel):
    """sample2D_linear (the bilinear sampler docs/09 asks for): texel centres at (i + 0.5) / size, repeat wrap."""
    info_t, kernel = linear_kernel
    rng = np.random.default_rng(5)
    tw, th = 7, 5
    tex = rng.random((th, tw, 4), dtype=np.float32)
    mem, desc = ren.create_texture2D(tw, th)
    with ren.mapped(mem) as m:
        m.view(np.float32).ravel().reshape(th, tw, 4)[:] = tex
    info = ren.create_struct(info_t)
    with ren.mapped(info) as m:
        m["Tex"] = desc.get()
    n = 4096
    uv = rng.uniform(-2.5, 2.5, (n, 2)).astype(np.float32)
    uv[:tw * th, 0] = (np.arange(tw * th) % tw + 0.5) / tw                   # texel centres: linear == nearest == the texel
    uv[:tw * th, 1] = (np.arange(tw * th) // tw + 0.5) / th
    cb = ren.create_buffer(n, ren.float2)
    with ren.mapped(cb) as m:
        m.view(np.float32).reshape(n, 2)[:] = uv
    near, lin = ren.create_buffer(n, ren.float4), ren.create_buffer(n, ren.float4)
    kernel[n](near, lin, cb, info)
    near, lin = near.get().view(np.float32).reshape(n, 4), lin.get().view(np.float32).reshape(n, 4)
    f32 = np.float32
    wrap = lambda c: np.fmod(np.fmod(c, f32(1)) + f32(1), f32(1)).astype(f32)
    x, y = wrap(uv[:, 0]) * f32(tw) - f32(0.5), wrap(uv[:, 1]) * f32(th) - f32(0.5)
    fx, fy = np.floor(x), np.floor(y)
    ax, ay = (x - fx).astype(f32)[:, None], (y - fy).astype(f32)[:, None]
    c0, r0 = fx.astype(int) % tw, fy.astype(int) % th
    c1, r1 = (c0 + 1) % tw, (r0 + 1) % th
    top = tex[r0, c0] * (f32(1) - ax) + tex[r0, c1] * ax
    bottom = tex[r1, c0] * (f32(1) - ax) + tex[r1, c1] * ax
    want = top * (f32(1) - ay) + bottom * ay
    assert np.allclose(lin, want, atol=1e-6)
    k = tw * th
    assert np.allclose(lin[:k], tex.reshape(-1, 4), atol=1e-6) and np.array_equal(near[:k], tex.reshape(-1, 4))


def test_struct_with_uint_vectors_keeps_the_host_layout(ren):
    """uint2/uint3/uint4 struct fields: OpenCL's 16-byte uint3, not CUDA's 12-byte built-in (the static_assert in the
    generated program checks sizeof against the host dtype)."""
    @ren.kernel_struct
    class UIntBag:
        a: ren.uint3
        b: ren.uint2
        c: ren.uint4
        d: np.uint32

    @ren.kernel_main
    def uint_bag_sum(bags: [UIntBag], out: [np.uint32]):
        """
        UIntBag b = bags[thread_id];
        out[thread_id] = b.a.x + b.a.y + b.a.z + b.b.x + b.b.y + b.c.w + b.d;
        """
    from rendering import _dsl
    assert _dsl.compile_program(_dsl.program_source(uint_bag_sum))
