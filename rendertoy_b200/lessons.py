"""The reference tutorials' raster set-ups as reusable functions (tests, bench.py and smoke() drive the product
through exactly the calls a tutorial script makes).

Shader bodies are the OpenCL C a user writes in tutorials/lesson08_rasterization.py:36-62 and
tutorials/lesson09_texture_mapping.py:67-95 (comments dropped); `rendering.Raster` recognises them by
fingerprint and runs their native CUDA twins.
"""
import numpy as np


def build_lesson08(ren, render_target):
    """-> (raster, shader_globals) as lesson08 builds them (:19-76)."""

    @ren.kernel_struct
    class Transforms:
        World: ren.float4x4
        View: ren.float4x4
        Proj: ren.float4x4

    @ren.kernel_struct
    class Vertex_Out:
        proj: ren.float4
        C: ren.float3

    shader_globals = ren.create_struct(Transforms)

    @ren.kernel_function
    def transform_and_draw(vertex: ren.MeshVertex, info: Transforms) -> Vertex_Out:
        """
        float3 P = vertex.P;
        float d = max(0.2f, dot(vertex.N, normalize((float3)(1,1,1))));
        float3 C = (float3)(d,d,d);
        float4 H = (float4)(P.x, P.y, P.z, 1.0);
        H = mul(H, info.World);
        H = mul(H, info.View);
        H = mul(H, info.Proj);
        Vertex_Out o;
        o.proj = H;
        o.C = C;
        return o;
        """

    @ren.kernel_function
    def fragment_to_color(fragment: Vertex_Out, info: Transforms) -> ren.float4:
        """
        return (float4)(fragment.C.x, fragment.C.y, fragment.C.z, 1);
        """

    raster = ren.Raster(render_target, transform_and_draw, shader_globals, fragment_to_color, shader_globals)
    return raster, shader_globals


def build_lesson09(ren, render_target, texture_rgb):
    """-> (raster, vertex_globals, fragment_globals, texture_descriptor) as lesson09 builds them (:33-107).
    texture_rgb: (h, w, 3) uint8 image (the tutorial loads models/marble2.jpg with PIL)."""
    h, w = texture_rgb.shape[0], texture_rgb.shape[1]
    texture_memory, texture_descriptor = ren.create_texture2D(w, h)
    with ren.mapped(texture_memory) as map:
        map = map.view(np.float32).ravel().reshape(h, w, 4)
        map[:, :, 0:3] = texture_rgb / 255.0
        map[:, :, 3] = 1.0

    @ren.kernel_struct
    class Transforms:
        World: ren.float4x4
        View: ren.float4x4
        Proj: ren.float4x4

    @ren.kernel_struct
    class Materials:
        DiffuseMap: ren.Texture2D

    @ren.kernel_struct
    class Vertex_Out:
        proj: ren.float4
        L: ren.float3
        C: ren.float2

    vertex_shader_globals = ren.create_struct(Transforms)
    fragment_shader_globals = ren.create_struct(Materials)

    @ren.kernel_function
    def transform_and_draw(vertex: ren.MeshVertex, info: Transforms) -> Vertex_Out:
        """
        float3 P = vertex.P;
        float d = 0.2f + max(0.0f, dot(vertex.N, normalize((float3)(1,1,1))));
        float3 L = (float3)(d,d,d);
        float4 H = (float4)(P.x, P.y, P.z, 1.0);
        H = mul(H, info.World);
        H = mul(H, info.View);
        H = mul(H, info.Proj);
        Vertex_Out o;
        o.proj = H;
        o.L = L;
        o.C = vertex.P.xy * 2;
        return o;
        """

    @ren.kernel_function
    def fragment_to_color(fragment: Vertex_Out, info: Materials) -> ren.float4:
        """
        float3 diff = sample2D(info.DiffuseMap, fragment.C).xyz;
        return (float4)(diff * fragment.L, 1);
        """

    raster = ren.Raster(render_target, transform_and_draw, vertex_shader_globals, fragment_to_color, fragment_shader_globals)
    with ren.mapped(fragment_shader_globals) as map:
        map["DiffuseMap"] = texture_descriptor.get()
    return raster, vertex_shader_globals, fragment_shader_globals, texture_descriptor


def set_transforms(ren, shader_globals, world, view, proj):
    """The per-frame `with ren.mapped(shader_globals) as map:` block of the tutorials (lesson08:90-99)."""
    with ren.mapped(shader_globals) as map:
        map["World"] = world
        map["View"] = view
        map["Proj"] = proj


def globals_as_floats(shader_globals):
    """48 float32 (World, View, Proj) for the oracle."""
    g = shader_globals.get()
    return np.concatenate([np.asarray(g[n]).reshape(-1).view(np.float32)[:16] for n in ("World", "View", "Proj")])


def render_frame(ren, raster, vertex_buffer, index_buffer=None, depth_clear=1.0):
    """clear + clear + draw, the body of every tutorial frame (lesson08:101-105)."""
    ren.clear(raster.get_render_target())
    ren.clear(raster.get_depth_buffer(), depth_clear)
    raster.draw_triangles(vertex_buffer, index_buffer)
