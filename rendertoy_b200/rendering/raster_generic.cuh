// raster_generic.cuh -- Raster pipeline for USER vertex / fragment shaders (rendering/_raster.py:63-112, 152-205,
// 227-327, 399-437), compiled at run time by NVRTC after cl_prelude.cuh, the user's structs and shader functions.
// rendering/_raster.py (ours) defines, before including this text:
//     VIN_T  VOUT_T  VSG_T  FSG_T   vertex-in / vertex-out / vertex-globals / fragment-globals struct types
//     VS_FN  FS_FN                  the @kernel_function shaders:  VOUT_T VS_FN(VIN_T, VSG_T);  float4 FS_FN(VOUT_T, FSG_T)
// The first field of VOUT_T is the float4 clip-space position (reference convention, _raster.py:78,121,139,213); every
// other field must be made of floats (they are interpolated component-wise like the reference's interpolate2/3).
//
// Same arithmetic, same quirks and same 64-bit (depth, primitive) key as the hand-written lesson08/09 kernels in
// csrc/rt_raster.cu -- results are bit-identical for the same shaders -- but a simpler schedule: one thread per
// triangle walks small bboxes itself, the warp shares out large ones.  It is the general path, not the fast one.

#define G_NF ((int)(sizeof(VOUT_T) / 4)) // floats per vertex-out, position first
#define G_NO_PRIM 0xFFFFFFFFu

union GVout { VOUT_T v; float f[sizeof(VOUT_T) / 4]; __device__ GVout() {} };

struct GEdges { float a1, b1, c1, a2, b2, c2, a3, b3, c3; unsigned tle; };
struct GBBox { int startx, starty, nx, ny; };
struct GCell { float al1, al2, al3; bool inside; };

__device__ inline GEdges g_edges(float h1x, float h1y, float h2x, float h2y, float h3x, float h3y)
{
    GEdges e; // _raster.py:274-292
    e.a1 = h2y - h1y; e.b1 = h1x - h2x; e.c1 = h1x * (h1y - h2y) - h1y * (h1x - h2x);
    e.a2 = h3y - h2y; e.b2 = h2x - h3x; e.c2 = h2x * (h2y - h3y) - h2y * (h2x - h3x);
    e.a3 = h1y - h3y; e.b3 = h3x - h1x; e.c3 = h3x * (h3y - h1y) - h3y * (h3x - h1x);
    e.tle = (((h1y == h2y && h2x <= h1x) || h1y < h2y) ? 1u : 0u) | (((h2y == h3y && h3x <= h2x) || h2y < h3y) ? 2u : 0u) |
            (((h3y == h1y && h1x <= h3x) || h3y < h1y) ? 4u : 0u);
    return e;
}

__device__ inline GBBox g_bbox(float x1, float y1, float x2, float y2, float x3, float y3, int W, int H)
{
    int minx = (int)fminf(x1, fminf(x2, x3)), miny = (int)fminf(y1, fminf(y2, y3)); // :237-241
    int maxx = (int)fmaxf(x1, fmaxf(x2, x3)), maxy = (int)fmaxf(y1, fmaxf(y2, y3));
    long long sx = max(0, minx), sy = max(0, miny);
    long long ex = min((long long)(W - 1), 1ll + maxx), ey = min((long long)(H - 1), 1ll + maxy);
    long long nx = ex - sx + 1, ny = ey - sy + 1;
    GBBox b; b.startx = (int)sx; b.starty = (int)sy;
    if (nx <= 0 || ny <= 0 || nx * ny >= 64 * 64) { b.nx = 0; b.ny = 0; } else { b.nx = (int)nx; b.ny = (int)ny; } // :294
    return b;
}

__device__ inline GCell g_cell(const GEdges &e, int col, int row)
{
    const float eps = 1e-8f; // :290-292
    float px = (float)col + 0.5f, py = (float)row + 0.5f;
    float d1 = e.a1 * px + e.b1 * py + e.c1, d2 = e.a2 * px + e.b2 * py + e.c2, d3 = e.a3 * px + e.b3 * py + e.c3;
    float s = d1 + d2 + d3;
    GCell c; c.al3 = d1 / s; c.al1 = d2 / s; c.al2 = d3 / s;
    c.inside = c.al1 >= ((e.tle & 2u) ? 0.0f : eps) && c.al2 >= ((e.tle & 4u) ? 0.0f : eps) && c.al3 >= ((e.tle & 1u) ? 0.0f : eps);
    return c;
}

__device__ inline void g_dehomogenize(float *p, float half_w, float half_h)
{
    const float w = p[3]; // :126-129
    p[0] = p[0] / w; p[1] = p[1] / w; p[2] = p[2] / w;
    p[1] = p[1] * -1.0f; p[0] = p[0] + 1.0f; p[1] = p[1] + 1.0f;
    p[0] = p[0] * half_w; p[1] = p[1] * half_h;
}

__device__ inline void g_lerp(const float *a, const float *b, float alpha, float *o)
{
    const float om = 1.0f - alpha; // interpolate2, :25-40
    for (int i = 0; i < G_NF; ++i) o[i] = a[i] * om + b[i] * alpha;
}

// exact coverage + depth atomic for the cells [c0, c1) (step `stride`) of a primitive given by its record
__device__ inline void g_cover(const float *r, unsigned prim, int c0, int stride, unsigned long long *key, int W, int H)
{
    const float *h1 = r, *h2 = r + G_NF, *h3 = r + 2 * G_NF;
    const GBBox bb = g_bbox(h1[0], h1[1], h2[0], h2[1], h3[0], h3[1], W, H);
    const GEdges e = g_edges(h1[0], h1[1], h2[0], h2[1], h3[0], h3[1]);
    const int n = bb.nx * bb.ny;
    for (int c = c0; c < n; c += stride) {
        const int rr = c / bb.nx, col = bb.startx + (c - rr * bb.nx), row = bb.starty + rr;
        const GCell cl = g_cell(e, col, row);
        if (!cl.inside) continue;
        const float hx = h1[0] * cl.al1 + h2[0] * cl.al2 + h3[0] * cl.al3;
        const float hy = h1[1] * cl.al1 + h2[1] * cl.al2 + h3[1] * cl.al3;
        const float hz = h1[2] * cl.al1 + h2[2] * cl.al2 + h3[2] * cl.al3;
        if (hz < 0) continue; // DepthTest :85
        const int ix = (int)hx, iy = (int)hy; // :88-89
        if (ix < 0 || ix >= W || iy < 0 || iy >= H) continue;
        atomicMin(key + (size_t)iy * W + ix, ((unsigned long long)__float_as_uint(hz) << 32) | prim);
    }
}

// VertexProcess + TriangleAssembly + Dehomogenize + TriangleRaster + DepthTest.  rec: 2 primitives x 3 x G_NF floats per triangle.
extern "C" __global__ void g_raster_triangles(const VIN_T *vb, const int *ib, VSG_T vsg, unsigned long long *key, float *rec, int W, int H,
                                              int number_of_threads)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const float half_w = (float)W * 0.5f, half_h = (float)H * 0.5f;
    int ncells[2] = {0, 0};
    if (t < number_of_threads) {
        GVout v[3], l01, l12, l20;
        for (int k = 0; k < 3; ++k) v[k].v = VS_FN(vb[ib ? ib[3 * t + k] : 3 * t + k], vsg);
        const float z0 = v[0].f[2], z1 = v[1].f[2], z2 = v[2].f[2];
        const int clip = (z0 < 0 ? 1 : 0) | (z1 < 0 ? 2 : 0) | (z2 < 0 ? 4 : 0); // :163
        int nprim = clip == 7 ? 0 : ((clip == 1 || clip == 2 || clip == 4) ? 2 : 1);
        if (clip != 0 && clip != 7) {
            g_lerp(v[0].f, v[1].f, -z0 / (z1 - z0), l01.f);
            g_lerp(v[1].f, v[2].f, -z1 / (z2 - z1), l12.f);
            g_lerp(v[2].f, v[0].f, -z2 / (z0 - z2), l20.f);
        }
        for (int k = 0; k < nprim; ++k) {
            const float *s1, *s2, *s3; // :178-199
            if (k == 0) switch (clip) {
                case 0: s1 = v[0].f; s2 = v[1].f; s3 = v[2].f; break;
                case 1: s1 = l01.f;  s2 = v[1].f; s3 = v[2].f; break;
                case 2: s1 = v[0].f; s2 = l01.f;  s3 = l12.f;  break;
                case 3: s1 = l12.f;  s2 = v[2].f; s3 = l20.f;  break;
                case 4: s1 = v[0].f; s2 = v[1].f; s3 = l12.f;  break;
                case 5: s1 = l01.f;  s2 = v[1].f; s3 = l12.f;  break;
                default: s1 = v[0].f; s2 = l01.f; s3 = l20.f;  break;
            } else switch (clip) {
                case 1: s1 = l01.f;  s2 = v[2].f; s3 = l20.f;  break;
                case 2: s1 = v[0].f; s2 = l12.f;  s3 = v[2].f; break;
                default: s1 = v[0].f; s2 = l12.f; s3 = l20.f;  break; // 4
            }
            float *r = rec + (size_t)(2 * t + k) * 3 * G_NF;
            for (int i = 0; i < G_NF; ++i) { r[i] = s1[i]; r[G_NF + i] = s2[i]; r[2 * G_NF + i] = s3[i]; }
            g_dehomogenize(r, half_w, half_h); g_dehomogenize(r + G_NF, half_w, half_h); g_dehomogenize(r + 2 * G_NF, half_w, half_h);
            if (r[2] < 0) continue; // :236
            const float e1x = r[G_NF] - r[0], e1y = r[G_NF + 1] - r[1], e2x = r[2 * G_NF] - r[0], e2y = r[2 * G_NF + 1] - r[1];
            if (!((e1x * e2y - e1y * e2x) <= 0)) // :259-266 swap v2, v3
                for (int i = 0; i < G_NF; ++i) { const float tmp = r[G_NF + i]; r[G_NF + i] = r[2 * G_NF + i]; r[2 * G_NF + i] = tmp; }
            const GBBox bb = g_bbox(r[0], r[1], r[G_NF], r[G_NF + 1], r[2 * G_NF], r[2 * G_NF + 1], W, H);
            ncells[k] = bb.nx * bb.ny;
        }
    }
    // coverage: a thread walks its own small primitives; the warp shares out the large ones
    for (int k = 0; k < 2; ++k) {
        if (ncells[k] > 0 && ncells[k] <= 64) g_cover(rec + (size_t)(2 * t + k) * 3 * G_NF, (unsigned)(2 * t + k), 0, 1, key, W, H);
        __syncwarp(); // the records written above must be visible to the other lanes of the warp
        unsigned big = __ballot_sync(0xffffffffu, ncells[k] > 64);
        while (big) {
            const int src = __ffs(big) - 1;
            big &= big - 1;
            const int tt = __shfl_sync(0xffffffffu, t, src);
            g_cover(rec + (size_t)(2 * tt + k) * 3 * G_NF, (unsigned)(2 * tt + k), lane, 32, key, W, H);
        }
    }
}

// FragmentProcess for the winners of this draw (one thread per pixel)
extern "C" __global__ void g_resolve_triangles(unsigned long long *key, const float *rec, FSG_T fsg, unsigned *bgra, int W, int H, int do_clear,
                                               unsigned clear_px, int number_of_threads)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= number_of_threads) return;
    const unsigned long long k64 = key[p];
    const unsigned prim = (unsigned)k64;
    if (prim == G_NO_PRIM) { if (do_clear) bgra[p] = clear_px; return; }
    const unsigned zbits = (unsigned)(k64 >> 32);
    const float *h1 = rec + (size_t)prim * 3 * G_NF, *h2 = h1 + G_NF, *h3 = h2 + G_NF;
    const GEdges e = g_edges(h1[0], h1[1], h2[0], h2[1], h3[0], h3[1]);
    const int x = p % W, y = p / W;
    GCell cl = g_cell(e, x, y);
    {
        const float hx = h1[0] * cl.al1 + h2[0] * cl.al2 + h3[0] * cl.al3, hy = h1[1] * cl.al1 + h2[1] * cl.al2 + h3[1] * cl.al3;
        const float hz = h1[2] * cl.al1 + h2[2] * cl.al2 + h3[2] * cl.al3;
        if (!(cl.inside && (int)hx == x && (int)hy == y && __float_as_uint(hz) == zbits)) {
            // the winning fragment came from another loop cell: first cell (row-major) of this primitive landing here
            const GBBox bb = g_bbox(h1[0], h1[1], h2[0], h2[1], h3[0], h3[1], W, H);
            bool found = false;
            for (int c = 0; c < bb.nx * bb.ny && !found; ++c) {
                const int rr = c / bb.nx;
                const GCell t = g_cell(e, bb.startx + (c - rr * bb.nx), bb.starty + rr);
                if (!t.inside) continue;
                const float tx = h1[0] * t.al1 + h2[0] * t.al2 + h3[0] * t.al3, ty = h1[1] * t.al1 + h2[1] * t.al2 + h3[1] * t.al3;
                const float tz = h1[2] * t.al1 + h2[2] * t.al2 + h3[2] * t.al3;
                if (!(tz < 0) && (int)tx == x && (int)ty == y && __float_as_uint(tz) == zbits) { cl = t; found = true; }
            }
        }
    }
    const float q1 = cl.al1 / h1[3], q2 = cl.al2 / h2[3], q3 = cl.al3 / h3[3]; // :313-318
    const float qs = q1 + q2 + q3, beta2 = q2 / qs, beta3 = q3 / qs, w1 = 1.0f - beta2 - beta3;
    GVout frag;
    for (int i = 4; i < G_NF; ++i) frag.f[i] = h1[i] * w1 + h2[i] * beta2 + h3[i] * beta3;
    for (int i = 0; i < 4; ++i) frag.f[i] = h1[i] * cl.al1 + h2[i] * cl.al2 + h3[i] * cl.al3;
    const clf4 color = FS_FN(frag.v, fsg);
    const float z = __uint_as_float(zbits);
    if (!(z <= 0)) bgra[p] = cl_unorm8(color.z) | (cl_unorm8(color.y) << 8) | (cl_unorm8(color.x) << 16) | (cl_unorm8(color.w) << 24);
    else if (do_clear) bgra[p] = clear_px;
    key[p] = k64 | 0xFFFFFFFFull;
}

// draw_points: VertexProcess + PointAssembly + PointRaster + Dehomogenize + DepthTest; rec: G_NF floats per point
extern "C" __global__ void g_raster_points(const VIN_T *vb, const int *ib, VSG_T vsg, unsigned long long *key, float *rec, int W, int H,
                                           int number_of_threads)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= number_of_threads) return;
    GVout v;
    v.v = VS_FN(vb[ib ? ib[i] : i], vsg);
    if (v.f[2] < 0) return;
    if (v.f[0] < -v.f[3] || v.f[0] > v.f[3] || v.f[1] < -v.f[3] || v.f[1] > v.f[3]) return;
    g_dehomogenize(v.f, (float)W * 0.5f, (float)H * 0.5f);
    float *r = rec + (size_t)i * G_NF;
    for (int k = 0; k < G_NF; ++k) r[k] = v.f[k];
    if (v.f[2] < 0) return;
    const int ix = (int)v.f[0], iy = (int)v.f[1];
    if (ix < 0 || ix >= W || iy < 0 || iy >= H) return;
    atomicMin(key + (size_t)iy * W + ix, ((unsigned long long)__float_as_uint(v.f[2]) << 32) | (unsigned)i);
}

extern "C" __global__ void g_resolve_points(unsigned long long *key, const float *rec, FSG_T fsg, unsigned *bgra, int W, int H, int do_clear,
                                            unsigned clear_px, int number_of_threads)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= number_of_threads) return;
    const unsigned long long k64 = key[p];
    const unsigned prim = (unsigned)k64;
    if (prim == G_NO_PRIM) { if (do_clear) bgra[p] = clear_px; return; }
    GVout frag;
    for (int k = 0; k < G_NF; ++k) frag.f[k] = rec[(size_t)prim * G_NF + k];
    const clf4 color = FS_FN(frag.v, fsg);
    const float z = __uint_as_float((unsigned)(k64 >> 32));
    if (!(z <= 0)) bgra[p] = cl_unorm8(color.z) | (cl_unorm8(color.y) << 8) | (cl_unorm8(color.x) << 16) | (cl_unorm8(color.w) << 24);
    else if (do_clear) bgra[p] = clear_px;
    key[p] = k64 | 0xFFFFFFFFull;
}
