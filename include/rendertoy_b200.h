/*
 * rendertoy_b200.h -- C ABI of librendertoy_b200.so (hand-written sm_100a CUDA behind the reference's
 * `rendering` Python package).  Plain pointers and sizes only; no torch / C++ types cross this line.
 *
 * The reference (lleonart1984/rendertoy) has no FFI of its own: its device boundary is pyopencl kernel
 * launches issued from rendering/_raster.py and rendering/_core.py.  Each entry point below names the
 * reference code it replaces (paths relative to the reference root).  INTEGRATION.md shows the ctypes
 * binding a maintainer adds to rendering/_raster.py / _raycaster.py to switch over.
 *
 * Conventions
 *   - every function returns 0 on success or a negative rt_status; rt_last_error() gives the text
 *     (thread-local, valid until the next failing call on that thread);
 *   - every `d_` pointer is DEVICE memory owned by the caller (the Python side passes
 *     torch.Tensor.data_ptr()); the library allocates nothing except texture objects and the NCCL
 *     communicator, which have explicit create/destroy calls;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream); all work is enqueued
 *     asynchronously on it and no entry point synchronises unless its comment says so;
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with RT_ERR_CUDA.
 */
#ifndef RENDERTOY_B200_H
#define RENDERTOY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RT_ABI_VERSION 1

typedef enum rt_status {
    RT_OK = 0,
    RT_ERR_INVALID = -1, /* bad argument */
    RT_ERR_CUDA = -2,    /* CUDA runtime error (text in rt_last_error) */
    RT_ERR_NCCL = -3,
    RT_ERR_UNSUPPORTED = -4
} rt_status;

/* Built-in shader pairs (vertex + fragment), recognised by the Python side from the tutorial sources. */
#define RT_SHADER_LESSON08 8 /* tutorials/lesson08_rasterization.py:36-62: Lambert max(0.2, N.l), colour = C      */
#define RT_SHADER_LESSON09 9 /* tutorials/lesson09_texture_mapping.py:67-95: L = 0.2+max(0,N.l), tex(P.xy*2) * L */

#define RT_NO_PRIMITIVE 0xFFFFFFFFu /* low word of a key no primitive has won in the current draw */

int rt_abi_version(void);
const char *rt_last_error(void);
/* SM count, L2 bytes, compute capability of the current device. */
int rt_device_info(int *sm_count, int *l2_bytes, int *cc_major, int *cc_minor);

/* ---- mesh upload -------------------------------------------------------------------------------
 * AoS MeshVertex[n] (80 B: P@0 N@16 C@32 T@48 B@64; rendering/_modeling.py:22-28, filled by
 * rendering/_loaders.py:16-38) -> SoA float4 position and normal arrays, the layout every kernel
 * below reads with 128-bit loads.  Replaces the per-draw 80 B/vertex reads of VertexProcess
 * (rendering/_raster.py:63-73). */
int rt_mesh_upload_soa(const void *d_mesh_vertices, int64_t n_vertices, void *d_pos4, void *d_nrm4, void *stream);

/* ---- render-target / depth clears  (rendering/_core.py:376-388 clear()) --------------------------
 * The depth buffer lives in the high word of a W*H array of 64-bit keys (depth_bits << 32 | primitive);
 * clearing writes (depth_bits << 32 | RT_NO_PRIMITIVE). */
int rt_raster_clear_depth(void *d_key, int64_t n_pixels, uint32_t depth_bits, void *stream);
/* BGRA8 target (CL_BGRA / CL_UNORM_INT8, rendering/_core.py:340): fill with rgba converted sat+rte. */
int rt_raster_clear_color(void *d_bgra, int64_t n_pixels, const float rgba[4], void *stream);
/* Raster.get_depth_buffer() view (rendering/_raster.py:388-389): gather / scatter the uint32 depth words. */
int rt_raster_read_depth(const void *d_key, int64_t n_pixels, void *d_depth_u32, void *stream);
int rt_raster_write_depth(void *d_key, int64_t n_pixels, const void *d_depth_u32, void *stream);

/* ---- Raster.draw_triangles  (rendering/_raster.py:416-437) -----------------------------------------
 * One call = VertexProcess + TriangleAssembly(+near clip) + Dehomogenize + TriangleRaster + DepthTest
 * + FragmentProcess of the reference (kernels at _raster.py:63-73, 152-205, 118-133, 227-327, 80-93,
 * 95-112), executed as three kernels: (1) fused vertex/clip/setup + coverage of small primitives with a
 * 64-bit atomicMin on the packed key, (1b) coverage of large primitives from a device work queue, (2) resolve:
 * re-interpolate the winning primitive per pixel, run the fragment shader once, write BGRA8, re-arm the key's
 * low word.  No host synchronisation, no intermediate fragment stream.
 *
 *   d_pos4, d_nrm4   SoA vertex arrays from rt_mesh_upload_soa
 *   d_indices        int32 triangle indices or NULL for a triangle soup (_raster.py:156-158)
 *   vs_globals       48 floats: Transforms{World, View, Proj} row-major (lesson08:19-23), HOST memory,
 *                    passed by value like the reference passes struct arguments (_core.py:270-274)
 *   tex_handle       0, or a handle from rt_texture_create (lesson09 Materials.DiffuseMap)
 *   d_key            W*H 64-bit keys (persistent across draws of a frame, like the depth buffer)
 *   d_scratch        >= rt_raster_scratch_bytes(...) bytes, 16-byte aligned, ZERO-FILLED once by the caller
 *                    (the first 256 bytes are a control block the kernels keep at zero between draws);
 *                    holds the per-primitive setup records and the large-primitive work queue
 *   d_bgra           W*H*4 bytes, the render target
 *   clear_rgba       NULL, or 4 floats: a clear(render_target, rgba) the caller deferred; folded into the resolve kernel
 *                    (pixels no primitive wins are written with it) instead of a separate fill launch
 *   clear_depth      non-zero: a deferred clear(depth_buffer, v) with v's bits in clear_depth_bits; executed first
 *   owner            NULL (the whole frame), or 7 ints {x0, y0, x1, y1, stripe_rows, stripe_mod, stripe_rem}: the pixels this
 *                    call may touch -- the inclusive rect and, inside it, the row stripes s = y / stripe_rows with
 *                    s % stripe_mod == stripe_rem (stripe_mod = 1: a plain scissor rect).  This is how one frame is split
 *                    over GPUs by image-space tiles (the reference is single-device, rendering/_core.py:10-11):
 *                    every rank draws the whole mesh with its own `owner`; owned pixels (key, depth, colour, folded clears)
 *                    end up exactly as without `owner`, all other pixels are not touched.  Primitives that cannot produce a
 *                    fragment in an owned pixel are dropped at setup, before any coverage work.
 */
int64_t rt_raster_scratch_bytes(int shader, int64_t n_triangles, int width, int height);
int rt_raster_draw_triangles(const void *d_pos4, const void *d_nrm4, const int32_t *d_indices, int64_t n_triangles,
                             int shader, const float *vs_globals, uint64_t tex_handle, int width, int height,
                             void *d_key, void *d_scratch, int64_t scratch_bytes, void *d_bgra, const float *clear_rgba,
                             int clear_depth, uint32_t clear_depth_bits, const int32_t *owner, void *stream);

/* ---- Raster.draw_points  (rendering/_raster.py:399-414) -----------------------------------------------------
 * VertexProcess + PointAssembly (z<0 cull, :140-151) + PointRaster (clip-space |x|,|y| <= w, :214-226) + Dehomogenize +
 * DepthTest + FragmentProcess as two kernels (per-point depth atomics, per-pixel resolve).  Arguments as for
 * rt_raster_draw_triangles; d_indices (int32 or NULL) selects the vertex of each point; point id = list position. */
int64_t rt_raster_points_scratch_bytes(int64_t n_points);
int rt_raster_draw_points(const void *d_pos4, const void *d_nrm4, const int32_t *d_indices, int64_t n_points, int shader,
                          const float *vs_globals, uint64_t tex_handle, int width, int height, void *d_key,
                          void *d_scratch, int64_t scratch_bytes, void *d_bgra, const float *clear_rgba, int clear_depth,
                          uint32_t clear_depth_bits, const int32_t *owner, void *stream);

/* ---- textures  (rendering/_core.py:551-578 MemoryPool / create_texture2D, :94-96 sample2D) ---------
 * Point-sampled float4 CUDA texture object over caller-owned linear device memory (row 0 first).
 * d_texels must be 512-byte aligned. */
int rt_texture_create(const void *d_texels, int width, int height, uint64_t *out_handle);
int rt_texture_destroy(uint64_t handle);

/* ---- ray casting  (rendering/_raycaster.py:8-36: BVH_AABB, BVH_Triangle, Raycaster) ----------------
 * The reference ships only the skeleton (ray_cast is `pass`); semantics are defined in
 * oracle/raycast_oracle.c.  Build: 30-bit Morton codes of triangle centroids -> LSD radix sort -> either the
 * Karras hierarchy + bottom-up AABB refit (RT_BVH_LBVH) or parallel locally-ordered clustering over the sorted
 * leaves (RT_BVH_PLOC: a few times slower to build, ~20 % fewer node visits per ray; falls back to the Karras tree
 * if the clustered tree came out deeper than the traversal's 64-entry stack).  Any correct tree gives the same
 * hits: the closest hit is defined by the exact triangle test alone. */
#define RT_BVH_LBVH 0
#define RT_BVH_PLOC 1
int64_t rt_bvh_node_bytes(int64_t n_triangles);    /* bytes of d_nodes   (64 B inner nodes)            */
int64_t rt_bvh_tri_bytes(int64_t n_triangles);     /* bytes of d_tris    (48 B leaf triangles, sorted) */
int64_t rt_bvh_scratch_bytes(int64_t n_triangles); /* bytes of d_scratch (build only)                  */
/* Raycaster._build_ads (rendering/_raycaster.py:30-33).  d_pos4 as for the rasterizer; triangle ids are
 * positions in the (optionally indexed) triangle list.  RT_BVH_PLOC synchronises the stream once (it reads the
 * finished tree's height back). */
int rt_bvh_build(const void *d_pos4, const int32_t *d_indices, int64_t n_triangles, void *d_nodes, void *d_tris,
                 void *d_scratch, int builder, void *stream);
/* Raycaster.ray_cast (rendering/_raycaster.py:35-36): rays = n x {float3 origin, float3 dir} (32 B, OpenCL
 * float3 padding), hits = n x {float t, uint32 triangle, float u, float v} (16 B); miss: t = +inf,
 * triangle = 0xFFFFFFFF. */
int rt_raycast_rays(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_rays, int64_t n_rays,
                    void *d_hits, void *stream);
/* Fused primary-ray generation + closest hit + Lambert/texture shade for the pixel rect
 * [x0,x0+w) x [y0,y0+h) of a width x height frame.  camera = {origin, U, V, W} (12 floats, model space,
 * HOST memory): dir = (U*sx + V*sy) + W with (sx, sy) the NDC pixel centre.  Outputs are rect-local,
 * row-major: d_hits (16 B/pixel, may be NULL), d_bgra (4 B/pixel, may be NULL; row pitch `bgra_pitch_px`
 * pixels, so a rank can write its tile straight into a full frame).  shader selects the lesson08 / lesson09
 * shading; d_pos4 / tex_handle are only needed for lesson09.  d_stats: NULL, or 3 x uint64 that an instrumented
 * build of the kernel ADDS {inner-node visits, triangle tests, rays} to (for the roofline report; slower).
 * cull_rect: NULL, or 4 ints {x0, y0, x1, y1} (inclusive, frame pixels): a conservative screen-space bound of the
 * scene the caller computed; pixels outside are written as misses without tracing.  fast_slab: non-zero lets the
 * box tests use the FMA form (caller guarantees the origin is within 16 scene extents of the scene).
 * d_view_nodes: NULL, or rt_raycast_view_node_bytes(n_triangles) of 16-byte aligned device scratch: the call then
 * first projects every BVH node into this camera's screen space (one small kernel) and the traversal tests pixels
 * against screen rectangles instead of rays against boxes -- same hits, fewer instructions per node.  The scratch is
 * per call: concurrent calls on different streams need different buffers.
 * stripes: NULL, or 3 ints {rows, mod, rem} (rows a multiple of 8, y0 a multiple of 8): the image-space partition of one
 * frame over `mod` GPUs -- frame rows are grouped from y = 0 into stripes of `rows` rows and only the stripes
 * s with s % mod == rem are traced, cleared and written; every other pixel of the rect (d_hits and d_bgra) is left untouched. */
int64_t rt_raycast_view_node_bytes(int64_t n_triangles);
/* After the projection, run `passes` (0..64) in-place passes that tighten every inner child's screen rectangle and depth
 * bound to the union of that child's own two (rt_raycast.cu: view_refit_kernel).  Hits cannot change (tested bit for bit);
 * node visits drop 11-26 % with 4 passes, each pass costs one more small launch.  Process-wide setting, default 0. */
int rt_raycast_set_view_refit(int passes);
/* Host only, no device work: the cull_rect for rt_raycast_primary -- conservative inclusive pixel rect of the scene box
 * [lo, hi] (3 doubles each) under `camera`, projected corners +- 2 px clamped to the frame.  Returns 1 and fills rect[4],
 * or 0 when there is no usable bound (box reaches the eye plane, singular camera basis, non-finite data): pass NULL then. */
/* Host only, no device work: the 12-float `camera` of rt_raycast_primary ({origin, U, V, W}, model space) from the
 * reference's View / Proj / World matrices (16 floats each, row-major, row-vector convention of rendering/_core.py:528-548;
 * world16 may be NULL).  Returns 1, or 0 for a singular view rotation / world matrix or non-finite data. */
int rt_camera_frame(const float *view16, const float *proj16, const float *world16, float *out12);
int rt_raycast_screen_bounds(const float *camera, const double *lo, const double *hi, int width, int height, int *rect);
/* Host only: the same for n boxes (n x 3 doubles each) -- the union of their rectangles; 0 as soon as one box has no bound.  With
 * the bounds of 64 chunks of the mesh instead of its one box the rectangle (= what a sparse read-back or gather moves, and
 * what is traced) is a quarter smaller for the dragon orbit. */
int rt_raycast_screen_bounds_n(const float *camera, const double *lo, const double *hi, int n_boxes, int width, int height, int *rect);
int rt_raycast_primary(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_pos4,
                       const void *d_nrm4, const int32_t *d_indices, const float *camera, int width, int height,
                       int x0, int y0, int w, int h, int shader, uint64_t tex_handle, void *d_hits, void *d_bgra,
                       int64_t bgra_pitch_px, void *d_stats, const int *cull_rect, int fast_slab,
                       void *d_view_nodes, const int32_t *stripes, void *stream);

/* ---- run-time kernels  (rendering/_core.py:247-299: kernel_main / build_kernel_main, one OpenCL program built at
 * first dispatch) --------------------------------------------------------------------------------------------
 * cuda_source: CUDA C++ produced from the user's OpenCL C by rendering/_dsl.py.  Compiled with NVRTC for sm_100a
 * (no FMA contraction); compile works without a GPU, launch needs one.  On a build error the compiler log is copied
 * to `log`.  rt_dsl_launch runs `kernel` over n_threads work-items, 1-D; args[i] points at the i-th argument's value
 * (device pointer for buffers, raw bytes for by-value structs/scalars), the trailing `int number_of_threads` included. */
int rt_dsl_compile(const char *cuda_source, uint64_t *out_module, char *log, int log_cap);
int rt_dsl_launch(uint64_t module, const char *kernel, int64_t n_threads, void **args, void *stream);
int rt_dsl_unload(uint64_t module);

/* Host only, no device work: where a draw with the tutorial vertex shaders (H = (((P, 1) World) View) Proj) can write.
 * globals48 = World, View, Proj (the Transforms struct); [lo, hi] = bounding box of the mesh (3 doubles each).  Returns 1
 * and the inclusive pixel rect of the eight projected corners +- 2 px clamped to the frame (x1 < x0: nothing on screen), or
 * 0 when the box reaches the near plane (triangles get clipped) or the data is not finite.  A frame is the clear colour
 * outside the union of its draws' rects: only that part has to be read back (rt_copy_rect). */
int rt_raster_screen_bounds(const float *globals48, const double *lo, const double *hi, int width, int height, int *rect);
int rt_raster_screen_bounds_n(const float *globals48, const double *lo, const double *hi, int n_boxes, int width, int height, int *rect);

/* ---- OBJ loading  (rendering/_loaders.py:7-33: pywavefront.Wavefront(path, collect_faces=True), then per mesh the first
 * material's interleaved, face-corner-expanded vertices) -- host code, no device work -----------------------------------
 * rt_obj_load parses the file (v / vn / vt / o / usemtl / f with v, v/t, v//n, v/t/n corners, 1-based or negative indices,
 * polygons fan-triangulated) and returns a handle; meshes are those that own at least one material, in file order.
 * rt_obj_mesh_info: soup vertex count (3 per triangle), triangle count and the vertex format (bit 0: T2F, bit 1: N3F; V3F
 * always) of mesh `mesh`.  rt_obj_mesh_rows scatters the soup into caller memory, `row_floats` floats per vertex (20 for
 * MeshVertex): P at [0..2], N at [4..6] when the format has normals, UV at [8..9] when it has texture coordinates; other
 * floats are left untouched.  Positions are as in the file: load_obj's normalisation (_loaders.py:34-38) is the caller's. */
int rt_obj_load(const char *path, uint64_t *out_handle);
int rt_obj_mesh_count(uint64_t handle);
int rt_obj_mesh_info(uint64_t handle, int mesh, int64_t *n_vertices, int64_t *n_faces, int *format);
int rt_obj_mesh_rows(uint64_t handle, int mesh, float *rows, int64_t row_floats);
int rt_obj_free(uint64_t handle);

/* ---- multi-GPU frame store  (no reference counterpart: rendering/_core.py:10-11 is single-device) -------
 * Rank 0 allocates the store (cudaMalloc, IPC-exportable) and exports a 64-byte handle; the other ranks of the
 * node open it and pass addresses inside it as `d_bgra` to rt_raycast_primary / rt_raster_draw_triangles /
 * rt_raster_clear_color, so finished pixels land in rank 0's HBM over NVLink from inside the shading kernel. */
int rt_peer_alloc(int64_t bytes, void **out_d_ptr);
int rt_peer_free(void *d_ptr);
int rt_peer_export(const void *d_ptr, void *handle64);
int rt_peer_open(const void *handle64, void **out_d_ptr);
int rt_peer_close(void *d_ptr);
/* The store starts cleared (rt_peer_alloc zero-fills it).  A ray-cast frame differs from the clear colour only inside the
 * scene's projected bounds, so a rank that rendered locally moves just that pixel rectangle (`rows` rows of `width_bytes`,
 * row pitches in bytes) into its slot of the store, or into a pinned host frame for the read-back (the reference side:
 * pyopencl enqueue_copy(queue, array, image, origin, region)): one pitched copy-engine transfer on `stream`; either
 * pointer may be local device, peer-mapped device or pinned host memory. */
int rt_copy_rect(void *d_dst, int64_t dst_pitch_bytes, const void *d_src, int64_t src_pitch_bytes, int64_t width_bytes,
                      int64_t rows, void *stream);
/* Sparse push by 32x32-pixel tiles, for frames that are mostly content (a frame-filling raster view: bounding rect ~90 % of
 * the frame, covered pixels a third): a kernel on the producing GPU stores only the tiles that hold a pixel != clear_px, or held
 * one the last time this producer pushed into the same destination (d_tile_state: rt_push_tiles_state_bytes(w, h) bytes of
 * producer-local device memory per destination frame, zero-filled when the destination is known to be all clear_px), into the
 * frame at d_dst (same W x H BGRA8 layout; rank 0's peer-mapped slot).  Afterwards the destination equals the source.  width
 * must be a multiple of 4.  d_bytes: NULL, or a device uint64 the kernel adds the bytes it stored to. */
int64_t rt_push_tiles_state_bytes(int width, int height);
int rt_push_tiles(void *d_dst, const void *d_src, int width, int height, uint32_t clear_px, void *d_tile_state, void *d_bytes,
                  void *stream);
/* The gather of the image-space partition: move the row stripes a rank owns -- rows y in [y0, y1] with
 * (y / stripe_rows) % mod == rem, bytes [x_bytes, x_bytes + width_bytes) of each row -- from the frame at d_src to the same
 * place in the frame at d_dst (same row pitch; local, peer-mapped or pinned host memory).  One 3-D copy-engine transfer for
 * the whole stripes plus at most two 2-D ones for stripes that [y0, y1] cuts. */
int rt_copy_stripes(void *d_dst, const void *d_src, int64_t pitch_bytes, int64_t x_bytes, int64_t width_bytes, int64_t y0, int64_t y1,
                    int stripe_rows, int mod, int rem, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* RENDERTOY_B200_H */
