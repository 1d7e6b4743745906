"""A small multi-mesh scene in the style of the reference's Class2022/*/scene.py, against the reference's OWN run of it
(tests/golden/scene2022.npz, made by oracle/clshim/make_golden.py: the unmodified reference package executed on the CPU).

What the scene exercises on top of the lesson tests: `manifold` grids with their int32 index buffers, @kernel_main kernels with
local arrays, for/break, rotation() and by-value scalar / matrix arguments (NVRTC DSL), a user shader pair with mesh UVs
(general raster path), several indexed draws with different textures composing on one depth/colour target, and a
draw_points overlay.  sin/cos differ by ulps between the host libm of the reference run and CUDA, so the kernels are compared
with a tolerance and the raster stage then starts from the reference's vertex data, where it has to be bit-exact."""
import os

import numpy as np
import pytest

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "scene2022.npz")
FIELDS = [0, 1, 2, 4, 5, 6, 8, 9, 12, 13, 14, 16, 17, 18]    # P, N, C, T, B of MeshVertex; the other floats are struct padding


def _scene(ren):
    z = np.load(GOLDEN)
    ns = {"ren": ren, "np": np}
    exec(str(z["source"]), ns)              # the very text the reference ran
    return z, ns


def test_manifold_grids_equal_the_reference_run(ren):
    """CPU: vertex grid and index buffer of ren.manifold against what the reference's manifold produced."""
    z, _ = _scene(ren)
    for i, (slices, stacks, _) in enumerate(z["specs"]):
        mesh = ren.manifold(int(slices), int(stacks))
        assert np.array_equal(mesh.vertices.get().view(np.float32).reshape(-1, 20)[:, FIELDS], z[f"grid{i}"][:, FIELDS])
        assert mesh.indices.dtype == np.int32 and np.array_equal(mesh.indices.get(), z[f"indices{i}"])


@pytest.mark.gpu
def test_scene_matches_reference_run(ren):
    z, ns = _scene(ren)
    w, h = int(z["width"]), int(z["height"])
    meshes = []
    for i, (slices, stacks, rscale) in enumerate(z["specs"]):
        mesh = ren.manifold(int(slices), int(stacks))
        if rscale > 0:
            ns["lathe"][mesh.vertices.shape](mesh.vertices, np.float32(rscale))
        got = mesh.vertices.get().view(np.float32).reshape(-1, 20)
        assert np.allclose(got[:, FIELDS], z[f"lathed{i}"][:, FIELDS], atol=3e-6), f"mesh {i}: lathe kernel differs from the reference run"
        ns["place"][mesh.vertices.shape](mesh.vertices, np.ascontiguousarray(z["place"][i]))
        got = mesh.vertices.get().view(np.float32).reshape(-1, 20)
        assert np.allclose(got[:, FIELDS], z[f"placed{i}"][:, FIELDS], atol=5e-6), f"mesh {i}: place kernel differs from the reference run"
        with ren.mapped(mesh.vertices) as m:                     # raster stage: from the reference's vertex data, bit for bit
            m.view(np.float32).reshape(-1, 20)[:] = z[f"placed{i}"]
        meshes.append(mesh)

    target = ren.create_image2d(w, h, ren._core.RGBA)
    g, fg = ren.create_struct(ns["SceneTransforms"]), ren.create_struct(ns["SceneMaterial"])
    raster = ren.Raster(target, ns["scene_vs"], g, ns["scene_fs"], fg)
    assert raster.shader_id is None, "a user shader pair must take the general (NVRTC) raster path"
    gl = z["globals"].reshape(3, 16)
    with ren.mapped(g) as m:
        m["World"], m["View"], m["Proj"] = (ren.make_float4x4(np.ascontiguousarray(x)) for x in gl)
    descs = []
    for i in range(len(meshes)):
        tex = z[f"texture{i}"]
        mem, desc = ren.create_texture2D(tex.shape[1], tex.shape[0])
        with ren.mapped(mem) as m:
            m = m.view(np.float32).ravel().reshape(tex.shape[0], tex.shape[1], 4)
            m[:, :, 0:3] = tex / 255.0
            m[:, :, 3] = 1.0
        descs.append(desc)
    ren.clear(raster.get_render_target())
    ren.clear(raster.get_depth_buffer(), 1.0)
    for mesh, desc in zip(meshes, descs):
        with ren.mapped(fg) as m:
            m["DiffuseMap"] = desc.get()
        raster.draw_triangles(mesh.vertices, mesh.indices)
    assert np.array_equal(raster.get_depth_buffer().get().reshape(h, w), z["depth_tris"]), "depth after the indexed draws differs"
    same = (target.get() == z["bgra_tris"]).all(axis=-1)
    assert same.mean() > 0.999, f"colour after the indexed draws: {same.mean():.5f} identical"   # depth ties are a race in the reference
    raster.draw_points(meshes[0].vertices)
    assert np.array_equal(raster.get_depth_buffer().get().reshape(h, w), z["depth"]), "depth after draw_points differs"
    same = (target.get() == z["bgra"]).all(axis=-1)
    print("scene2022: identical pixels", same.mean())
    assert same.mean() > 0.999
