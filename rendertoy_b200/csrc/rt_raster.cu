// rt_raster.cu -- Raster.draw_triangles / draw_points for sm_100a (replaces rendering/_raster.py:416-437 and the seven
// OpenCL kernels it drives).  One C-ABI call = up to four launches on the caller's stream, no host round trip, no
// fragment stream:
//
//   fill_u64        a pending clear(depth_buffer) folded into the draw: key = depth_bits << 32 | NO_PRIMITIVE
//   raster_kernel   one thread per input triangle: vertex shader x3, near clip into <=2 primitives, dehomogenize, bbox +
//                   edge setup (all bit-identical to the reference arithmetic), one 64 B/96 B record per primitive; then
//                   each warp stages its primitives in shared memory, finds the exact accepted interval of every bbox
//                   row, and runs the reference's coverage test + a 64-bit atomicMin of (depth_bits << 32 | primitive
//                   id) into the key buffer on full warps of candidate cells.  Large primitives go to a work queue.
//   coverage_kernel persistent grid draining that queue (row-quads of large primitives)
//   resolve_kernel  four pixels per thread: decode the winning key, re-evaluate that primitive at the pixel from its
//                   record, perspective-correct attributes, fragment shader once, BGRA8 store (or the folded clear
//                   colour), re-arm the key's low word for the next draw.
//
// `owner` (scissor rect + row stripes) restricts all of it to the pixels one rank owns (image-space partition).
// Bound: instruction issue (raster_kernel) and key -> record latency (resolve_kernel); HBM sees vertices in and
// keys / colour out at ~16 % of its roofline; no dense contraction, so no tensor cores.
#include "rt_common.cuh"

namespace {

struct VO { // vertex-shader output: clip-space position + up to 3 attribute floats
    float x, y, z, w, a0, a1, a2;
};

struct DrawArgs {
    const float4 *pos, *nrm;
    const int *idx;
    long long n_tris;
    float g[48]; // World, View, Proj
    int width, height;
    float half_w, half_h;
    unsigned long long *key;
    float4 *rec;
    RtOwner own; // pixels this draw may write ...
    int scissor; // ... 0: the whole frame (no ownership tests anywhere)
};

// ---- reference arithmetic, restated -----------------------------------------------------------

// _core.py:86-88 mul(float4, float4x4): r_j = dot(v, column j)
__device__ __forceinline__ float4 mul4(float4 v, const float *m)
{
    float4 r;
    r.x = ((v.x * m[0] + v.y * m[4]) + v.z * m[8]) + v.w * m[12];
    r.y = ((v.x * m[1] + v.y * m[5]) + v.z * m[9]) + v.w * m[13];
    r.z = ((v.x * m[2] + v.y * m[6]) + v.z * m[10]) + v.w * m[14];
    r.w = ((v.x * m[3] + v.y * m[7]) + v.z * m[11]) + v.w * m[15];
    return r;
}

// lesson08:41-54 (SHADER 8) / lesson09:72-86 (SHADER 9)
template <int SHADER>
__device__ __forceinline__ VO vertex_shader(float4 P, float4 N, const float *g)
{
    const float n = RT_INV_SQRT3;
    float dt = (N.x * n + N.y * n) + N.z * n;
    float4 H = make_float4(P.x, P.y, P.z, 1.0f);
    H = mul4(H, g);
    H = mul4(H, g + 16);
    H = mul4(H, g + 32);
    VO o;
    o.x = H.x; o.y = H.y; o.z = H.z; o.w = H.w;
    if (SHADER == RT_SHADER_LESSON08) {
        o.a0 = fmaxf(0.2f, dt); o.a1 = 0.0f; o.a2 = 0.0f;
    } else {
        o.a0 = 0.2f + fmaxf(0.0f, dt); o.a1 = P.x * 2.0f; o.a2 = P.y * 2.0f;
    }
    return o;
}

// _raster.py:25-40 interpolate2: v0*(1-alpha) + v1*alpha on every field
template <int SHADER>
__device__ __forceinline__ VO lerp_vo(const VO &a, const VO &b, float alpha)
{
    const float om = 1.0f - alpha;
    VO o;
    o.x = a.x * om + b.x * alpha; o.y = a.y * om + b.y * alpha;
    o.z = a.z * om + b.z * alpha; o.w = a.w * om + b.w * alpha;
    o.a0 = a.a0 * om + b.a0 * alpha;
    if (SHADER == RT_SHADER_LESSON09) { o.a1 = a.a1 * om + b.a1 * alpha; o.a2 = a.a2 * om + b.a2 * alpha; }
    else { o.a1 = 0.0f; o.a2 = 0.0f; }
    return o;
}

// _raster.py:126-129 Dehomogenize
__device__ __forceinline__ void dehomogenize(VO &p, float half_w, float half_h)
{
    p.x = p.x / p.w; p.y = p.y / p.w; p.z = p.z / p.w;
    p.y = p.y * -1.0f;
    p.x = p.x + 1.0f; p.y = p.y + 1.0f;
    p.x = p.x * half_w; p.y = p.y * half_h;
}

struct Edges { // _raster.py:274-292
    float a1, b1, c1, a2, b2, c2, a3, b3, c3;
    unsigned tle; // bit0: v1v2, bit1: v2v3, bit2: v3v1 is a top/left edge
};

__device__ __forceinline__ Edges edge_setup(float h1x, float h1y, float h2x, float h2y, float h3x, float h3y)
{
    Edges e;
    e.a1 = h2y - h1y; e.b1 = h1x - h2x; e.c1 = h1x * (h1y - h2y) - h1y * (h1x - h2x);
    e.a2 = h3y - h2y; e.b2 = h2x - h3x; e.c2 = h2x * (h2y - h3y) - h2y * (h2x - h3x);
    e.a3 = h1y - h3y; e.b3 = h3x - h1x; e.c3 = h3x * (h3y - h1y) - h3y * (h3x - h1x);
    unsigned t12 = ((h1y == h2y && h2x <= h1x) || h1y < h2y) ? 1u : 0u;
    unsigned t23 = ((h2y == h3y && h3x <= h2x) || h2y < h3y) ? 2u : 0u;
    unsigned t31 = ((h3y == h1y && h1x <= h3x) || h3y < h1y) ? 4u : 0u;
    e.tle = t12 | t23 | t31;
    return e;
}

struct BBox { int startx, starty, nx, ny; }; // nx*ny == 0 when nothing is to be rasterized

// _raster.py:237-241 + the `pixel_count < 64*64` gate of :294
__device__ __forceinline__ BBox bbox_setup(float x1, float y1, float x2, float y2, float x3, float y3, int W, int H)
{
    int minx = (int)fminf(x1, fminf(x2, x3)), miny = (int)fminf(y1, fminf(y2, y3));
    int maxx = (int)fmaxf(x1, fmaxf(x2, x3)), maxy = (int)fmaxf(y1, fmaxf(y2, y3));
    long long sx = max(0, minx), sy = max(0, miny);
    long long ex = min((long long)(W - 1), 1ll + maxx), ey = min((long long)(H - 1), 1ll + maxy);
    long long nx = ex - sx + 1, ny = ey - sy + 1;
    BBox b;
    b.startx = (int)sx; b.starty = (int)sy;
    if (nx <= 0 || ny <= 0 || nx * ny >= 64 * 64) { b.nx = 0; b.ny = 0; }
    else { b.nx = (int)nx; b.ny = (int)ny; }
    return b;
}

struct Cell { // _raster.py:298-311 for one (col,row)
    float al1, al2, al3;
    bool inside;
};

__device__ __forceinline__ Cell cell_eval(const Edges &e, int col, int row)
{
    const float eps = 1e-8f; // (float)0.00000001, _raster.py:290-292
    float px = (float)col + 0.5f, py = (float)row + 0.5f;
    float d1 = e.a1 * px + e.b1 * py + e.c1;
    float d2 = e.a2 * px + e.b2 * py + e.c2;
    float d3 = e.a3 * px + e.b3 * py + e.c3;
    float s = d1 + d2 + d3;
    Cell c;
    c.al3 = d1 / s; c.al1 = d2 / s; c.al2 = d3 / s;
    float comp3 = (e.tle & 1u) ? 0.0f : eps, comp1 = (e.tle & 2u) ? 0.0f : eps, comp2 = (e.tle & 4u) ? 0.0f : eps;
    c.inside = c.al1 >= comp1 && c.al2 >= comp2 && c.al3 >= comp3;
    return c;
}

__device__ __forceinline__ float blend3(float f1, float f2, float f3, float w1, float w2, float w3)
{
    return f1 * w1 + f2 * w2 + f3 * w3;
}

// ---- kernel 1: vertex + clip + setup + coverage + depth atomics ---------------------------------
//
// Warp-autonomous: every warp owns 32 consecutive input triangles and never talks to another warp (no
// __syncthreads).  After setup,
//   * primitives with a small bbox (<= SMALL_MAX cells) are compacted into the warp's 32 shared-memory slots,
//     their cells laid end to end, and the warp walks that list 32 cells at a time so lanes stay busy
//     whatever the mix of sizes.  A cell finds its primitive with one warp OR-reduce + popc: owners flag the
//     cell where their primitive starts, and a cell's slot is the number of starts at or before it;
//   * primitives with a large bbox are NOT walked here -- one warp stuck on a few thousand cells would be the
//     kernel's critical path.  They are cut into work items of 32 row-quads appended to a global queue
//     (one warp-aggregated atomicAdd) that coverage_kernel drains with every warp of the machine.
// Candidate cells (those the division-free sign test cannot reject) are compacted through a per-warp
// shared-memory ring, so the exact test + depth atomics always run with full warps.

#ifndef RT_RASTER_RW
#define RT_RASTER_RW 4
#endif
#ifndef RT_RASTER_MINB
// resident blocks per SM the compiler must allow for (register cap); -D overrides for A/B builds.  Measured on B200 (cfg2, 8 frame
// streams): uncapped (78 registers) 51.1 us per frame, 7 blocks (72 registers, no spills for lesson08) 49.6, 8 blocks 49.5,
// 2- or 8-warp blocks 49.4 / 49.7; resolve_kernel at 2 or 4 blocks per SM instead of 3: 53.2 / 53.5.
#define RT_RASTER_MINB (SHADER == RT_SHADER_LESSON08 && !SC ? 7 : 1)
#endif
#ifndef RT_COVERAGE_BLOCKS_PER_SM
// persistent coverage grid: blocks of 128 threads per SM (-D overrides for A/B builds).  Measured on B200 (cfg2, 8 frame streams):
// 16 blocks per SM 48.7 us per frame, 8 -> 48.4, 4 -> 47.8 (one frame alone +1 us), 2 -> 47.7 (alone +4 us): a smaller grid leaves
// room for the other frames' kernels while it waits for its few work items.
#define RT_COVERAGE_BLOCKS_PER_SM 4
#endif
#ifndef RT_RESOLVE_MINB
#define RT_RESOLVE_MINB 3
#endif
constexpr int RW = RT_RASTER_RW; // warps per raster block
constexpr int SMALL_MAX = 256;  // largest bbox (cells) rasterized inline by the owning warp
constexpr int QUADS = 32;       // a queued work item = 32 row-quads (4 cells along x each) of a large primitive
constexpr int RING = 160;       // per-warp candidate stack: < 32 left over + up to 4 new per lane per pass

struct __align__(16) Slot { // 24 words, read as six 128-bit loads (the last two only for candidate cells)
    float a1, b1, c1, a2;
    float b2, c2, a3, b3;
    float c3, inv_nx; int off; int sxy;        // sxy = startx | starty << 16
    int nx_tle; unsigned prim; float h1x, h1y; // nx_tle = nx | tle << 16 | robust << 19 | (s < 0) << 20 | local << 21 (scissor)
    float h1z, h2x, h2y, h2z;
    float h3x, h3y, h3z, pad;
};

struct WorkCtl { // head of the scratch buffer (zero-filled by the caller once, kept at zero between draws)
    unsigned n_items;    // queued work items of this draw (reset by resolve_kernel)
    unsigned overflowed; // warps that had to rasterize large primitives inline because the queue was full
    unsigned pad[2];
};

template <int SHADER> struct RecLayout { static constexpr int F4 = SHADER == RT_SHADER_LESSON08 ? 4 : 6; };

// Exact coverage test + depth atomic for one cell of a primitive (h = dehomogenized vertices in edge order).
template <bool SC>
__device__ __forceinline__ void cover_cell(const Edges &e, int col, int row, float h1x, float h1y, float h1z, float h2x,
                                           float h2y, float h2z, float h3x, float h3y, float h3z, unsigned prim,
                                           unsigned long long *key, int W, int H, const DrawArgs &a)
{
    Cell cl = cell_eval(e, col, row);
    if (!cl.inside) return;
    float hx = blend3(h1x, h2x, h3x, cl.al1, cl.al2, cl.al3);
    float hy = blend3(h1y, h2y, h3y, cl.al1, cl.al2, cl.al3);
    float hz = blend3(h1z, h2z, h3z, cl.al1, cl.al2, cl.al3);
    if (hz < 0) return; // DepthTest, _raster.py:85
    int ix = (int)hx, iy = (int)hy; // the INTERPOLATED position picks the pixel (:88-89)
    if (ix < 0 || ix >= W || iy < 0 || iy >= H) return;
    if (SC && !rt_owns(a.own, ix, iy)) return; // the exact ownership test: on the pixel the fragment lands on
    unsigned long long k64 = ((unsigned long long)__float_as_uint(hz) << 32) | prim;
    atomicMin(key + (size_t)iy * W + ix, k64);
}

// Row spans need sign(s) to be one constant over the whole bbox (s = d1 + d2 + d3, twice the signed area up to
// rounding).  In real arithmetic s is affine in (px, py), so it lies between its four corner values; the float
// evaluation differs from it by less than N/2 with N = 2^-20 * sum of term magnitudes (a generous 16 ulp).  Hence if
// all four corner values exceed N with one sign, every cell's float s has that sign and is non-zero.  Returns
// 0 (not robust: slivers, NaN/inf), +1 or -1.
__device__ __forceinline__ int robust_sign(const Edges &e, const BBox &bb, float &s_abs_min, float &noise)
{
    s_abs_min = 0.0f; noise = INFINITY;
    const float x0 = (float)bb.startx + 0.5f, x1 = (float)(bb.startx + bb.nx - 1) + 0.5f;
    const float y0 = (float)bb.starty + 0.5f, y1 = (float)(bb.starty + bb.ny - 1) + 0.5f;
    const float n = 9.5367431640625e-7f * ((fabsf(e.a1) + fabsf(e.a2) + fabsf(e.a3)) * x1 + (fabsf(e.b1) + fabsf(e.b2) + fabsf(e.b3)) * y1 +
                                           (fabsf(e.c1) + fabsf(e.c2) + fabsf(e.c3)));
    float smin = INFINITY, smax = -INFINITY;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float px = (c & 1) ? x1 : x0, py = (c & 2) ? y1 : y0;
        const float d1 = e.a1 * px + e.b1 * py + e.c1, d2 = e.a2 * px + e.b2 * py + e.c2, d3 = e.a3 * px + e.b3 * py + e.c3;
        const float sv = d1 + d2 + d3;
        smin = fminf(smin, sv); smax = fmaxf(smax, sv);
        if (!(sv == sv)) return 0;
    }
    if (!(n > 1e-18f) || !(n < INFINITY)) return 0;
    noise = n;
    if (smin > n) { s_abs_min = smin; return 1; }
    if (smax < -n) { s_abs_min = -smax; return -1; }
    return 0;
}

template <int SHADER, bool SC>
__device__ __forceinline__ int setup_prim(const DrawArgs &a, VO p1, VO p2, VO p3, unsigned prim, Slot &s)
{
    dehomogenize(p1, a.half_w, a.half_h);
    dehomogenize(p2, a.half_w, a.half_h);
    dehomogenize(p3, a.half_w, a.half_h);
    if (p1.z < 0) return 0; // _raster.py:236
    BBox bb = bbox_setup(p1.x, p1.y, p2.x, p2.y, p3.x, p3.y, a.width, a.height);
    if (bb.nx == 0) return 0;
    if (SC && !rt_owns_any_row(a.own, bb.starty - 1, bb.starty + bb.ny)) return 0; // cheap early out (re-checked below)
    float e1x = p2.x - p1.x, e1y = p2.y - p1.y, e2x = p3.x - p1.x, e2y = p3.y - p1.y;
    if (!((e1x * e2y - e1y * e2x) <= 0)) { VO t = p2; p2 = p3; p3 = t; } // :259-266
    Edges e = edge_setup(p1.x, p1.y, p2.x, p2.y, p3.x, p3.y);
    float s_abs_min, noise;
    const int rs = robust_sign(e, bb, s_abs_min, noise);
    int local = 0;
    if (SC) {
        // Ownership is decided per FRAGMENT (cover_cell: the pixel the interpolated position lands on, _raster.py:88-89), so a
        // cell may only be skipped when its fragment provably lands within a pixel of the cell.  The interpolated position is
        // px + sum_k (h_k - p)(alpha_k - lambda_k) + O(2^-22 px): the float barycentrics alpha_k = d_k / s are off their real
        // values lambda_k by at most 2 noise / |s| each (noise bounds the rounding of a float edge sum, see robust_sign), so the
        // fragment moves by less than 6 noise ext / |s|, ext = the triangle's extent.  With |s| > 8 noise ext that is < 0.75 px
        // and every cell outside the owned rect grown by one pixel can be dropped: such primitives are "local", their bbox
        // is clipped (after the 64x64 gate, which the reference evaluates on the screen-clamped bbox) and rows of stripes the
        // rank does not own are skipped.  Slivers that fail the bound keep their whole bbox; only cover_cell filters them.
        const float ext = fmaxf(fmaxf(p1.x, fmaxf(p2.x, p3.x)) - fminf(p1.x, fminf(p2.x, p3.x)),
                                fmaxf(p1.y, fmaxf(p2.y, p3.y)) - fminf(p1.y, fminf(p2.y, p3.y)));
        local = rs != 0 && s_abs_min > 8.0f * noise * ext;
        if (local) {
            const int cx0 = max(bb.startx, a.own.x0 - 1), cx1 = min(bb.startx + bb.nx - 1, a.own.x1 + 1);
            const int cy0 = max(bb.starty, a.own.y0 - 1), cy1 = min(bb.starty + bb.ny - 1, a.own.y1 + 1);
            if (cx1 < cx0 || cy1 < cy0 || !rt_owns_any_row(a.own, cy0 - 1, cy1 + 1)) return 0;
            bb.startx = cx0; bb.starty = cy0; bb.nx = cx1 - cx0 + 1; bb.ny = cy1 - cy0 + 1;
        }
    }

    float4 *r = a.rec + (size_t)prim * RecLayout<SHADER>::F4;
    r[0] = make_float4(p1.x, p1.y, p1.z, p1.w);
    r[1] = make_float4(p2.x, p2.y, p2.z, p2.w);
    r[2] = make_float4(p3.x, p3.y, p3.z, p3.w);
    r[3] = make_float4(p1.a0, p2.a0, p3.a0, 0.0f);
    if (SHADER == RT_SHADER_LESSON09) {
        r[4] = make_float4(p1.a1, p1.a2, p2.a1, p2.a2);
        r[5] = make_float4(p3.a1, p3.a2, 0.0f, 0.0f);
    }
    s.a1 = e.a1; s.b1 = e.b1; s.c1 = e.c1;
    s.a2 = e.a2; s.b2 = e.b2; s.c2 = e.c2;
    s.a3 = e.a3; s.b3 = e.b3; s.c3 = e.c3;
    s.inv_nx = 1.0f / (float)bb.nx;
    s.sxy = bb.startx | (bb.starty << 16);
    // nx | tle<<16 | robust<<19 | s<0 <<20 | local<<21
    s.nx_tle = bb.nx | ((int)e.tle << 16) | (rs != 0 ? 1 << 19 : 0) | (rs < 0 ? 1 << 20 : 0) | (local ? 1 << 21 : 0);
    s.prim = prim;
    s.h1x = p1.x; s.h1y = p1.y; s.h1z = p1.z;
    s.h2x = p2.x; s.h2y = p2.y; s.h2z = p2.z;
    s.h3x = p3.x; s.h3y = p3.y; s.h3z = p3.z;
    s.pad = 0.0f;
    s.off = bb.ny; // large primitives keep ny here; the inline path overwrites it with the cell offset
    return bb.nx * bb.ny;
}

// _raster.py:152-205 TriangleAssembly: clip code, lerped vertices, first/second output triangle
template <int SHADER>
__device__ __forceinline__ int assemble(const DrawArgs &a, long long t, VO (&q)[2][3])
{
    long long i0 = 3 * t, i1 = 3 * t + 1, i2 = 3 * t + 2;
    if (a.idx) { i0 = a.idx[i0]; i1 = a.idx[i1]; i2 = a.idx[i2]; }
    VO v0 = vertex_shader<SHADER>(__ldg(a.pos + i0), __ldg(a.nrm + i0), a.g);
    VO v1 = vertex_shader<SHADER>(__ldg(a.pos + i1), __ldg(a.nrm + i1), a.g);
    VO v2 = vertex_shader<SHADER>(__ldg(a.pos + i2), __ldg(a.nrm + i2), a.g);
    int clip = (v0.z < 0 ? 1 : 0) | (v1.z < 0 ? 2 : 0) | (v2.z < 0 ? 4 : 0);
    if (clip == 0) { q[0][0] = v0; q[0][1] = v1; q[0][2] = v2; return 1; }
    if (clip == 7) return 0;
    VO v01 = lerp_vo<SHADER>(v0, v1, -v0.z / (v1.z - v0.z));
    VO v12 = lerp_vo<SHADER>(v1, v2, -v1.z / (v2.z - v1.z));
    VO v20 = lerp_vo<SHADER>(v2, v0, -v2.z / (v0.z - v2.z));
    switch (clip) {
    case 1: q[0][0] = v01; q[0][1] = v1;  q[0][2] = v2;  q[1][0] = v01; q[1][1] = v2;  q[1][2] = v20; return 2;
    case 2: q[0][0] = v0;  q[0][1] = v01; q[0][2] = v12; q[1][0] = v0;  q[1][1] = v12; q[1][2] = v2;  return 2;
    case 3: q[0][0] = v12; q[0][1] = v2;  q[0][2] = v20; return 1;
    case 4: q[0][0] = v0;  q[0][1] = v1;  q[0][2] = v12; q[1][0] = v0;  q[1][1] = v12; q[1][2] = v20; return 2;
    case 5: q[0][0] = v01; q[0][1] = v1;  q[0][2] = v12; return 1;
    default: q[0][0] = v0; q[0][1] = v01; q[0][2] = v20; return 1; // 6
    }
}

// Exact test for one ring entry (slot << 24 | row_local << 12 | col_local packed by the producer).
template <bool SC>
__device__ __forceinline__ void cover_from_slot(const Slot *my_slots, unsigned packed, unsigned long long *key, int W, int H, const DrawArgs &a)
{
    const float4 *sp = reinterpret_cast<const float4 *>(my_slots + (packed >> 24));
    const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2], s3 = sp[3], s4 = sp[4], s5 = sp[5];
    const int sxy = __float_as_int(s2.w), nx_tle = __float_as_int(s3.x);
    const int col = (sxy & 0xffff) + (int)(packed & 0xfffu), row = (sxy >> 16) + (int)((packed >> 12) & 0xfffu);
    Edges e;
    e.a1 = s0.x; e.b1 = s0.y; e.c1 = s0.z; e.a2 = s0.w; e.b2 = s1.x; e.c2 = s1.y;
    e.a3 = s1.z; e.b3 = s1.w; e.c3 = s2.x; e.tle = (unsigned)(nx_tle >> 16) & 7u;
    cover_cell<SC>(e, col, row, s3.z, s3.w, s4.x, s4.y, s4.z, s4.w, s5.x, s5.y, s5.z, __float_as_uint(s3.y), key, W, H, a);
}

// Coverage of the primitives staged in `my_slots` (compacted, owner lane i holds ny / off of slot i): their rows are
// laid end to end and taken 32 at a time, one row per lane.  A lane finds the exact interval of its row that the
// division-free sign test accepts (the float edge functions are monotone in px, so per edge the accepted cells are a
// half-line: an analytic guess is corrected by evaluating the real predicate), pushes those cells on the warp's
// candidate stack, and full warps of candidates go through the exact test + depth atomics.
template <bool SC>
__device__ __forceinline__ void cover_rows(const Slot *my_slots, unsigned *my_ring, int ny, int off, int total, int lane,
                                           unsigned long long *key, int W, int H, const DrawArgs &a)
{
    const unsigned FULL = 0xffffffffu, le_mask = (2u << lane) - 1u;
    int started = 0; // compacted primitives whose first row lies before `base`
    int pending = 0; // candidates on the stack (warp-uniform, < 32 between passes)
    for (int base = 0; base < total; base += 32) {
        const unsigned rel = (unsigned)(off - base);
        const unsigned starts = __reduce_or_sync(FULL, (ny > 0 && rel < 32u) ? (1u << rel) : 0u);
        const int slot = started + __popc(starts & le_mask) - 1;
        started += __popc(starts);
        int lo = 0, hi = -1, lrow = 0;
        if (base + lane < total) {
            const float4 *sp = reinterpret_cast<const float4 *>(my_slots + slot);
            const float4 s0 = sp[0], s1 = sp[1], s2 = sp[2];
            const int nx_tle = __float_as_int(sp[3].x), sxy = __float_as_int(s2.w);
            lrow = base + lane - __float_as_int(s2.z);
            hi = (nx_tle & 0xffff) - 1;
            // a row of a "local" primitive (setup_prim) with no owned row within one pixel produces no owned fragment
            if (SC && (nx_tle & (1 << 21)) && !rt_owns_any_row(a.own, (sxy >> 16) + lrow - 1, (sxy >> 16) + lrow + 1)) hi = -1;
            if (nx_tle & (1 << 19)) { // sign(s) is one constant over the bbox: exact spans
                const float sg = (nx_tle & (1 << 20)) ? -1.0f : 1.0f;
                const float x0 = (float)(sxy & 0xffff) + 0.5f; // px of local column 0
                const float py = (float)((sxy >> 16) + lrow) + 0.5f;
                const float ea[3] = {s0.x, s0.w, s1.z}, eb[3] = {s0.y, s1.x, s1.w}, ec[3] = {s0.z, s1.y, s2.x};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float ak = ea[k], tk = eb[k] * py, ck = ec[k];
                    // accepted(col) <=> !(sg * ((ak * px + tk) + ck) < 0), px = x0 + col  (same expression as cell_eval)
                    const float dir = sg * ak;
                    if (dir > 0.0f || dir < 0.0f) {
                        // accepted cells are a half-line: col >= boundary (dir > 0, step +1) or col <= boundary (dir < 0,
                        // step -1).  One code path for both directions, so lanes of differently oriented edges do not
                        // diverge: guess analytically, walk outwards while the cell before is still accepted, then
                        // inwards while the cell itself is rejected -- the real predicate decides, the guess only seeds.
                        if (hi >= lo) {
                            const int st = dir > 0.0f ? 1 : -1;
                            const float q = __fdividef(-(tk + ck), ak) - x0;
                            int g = st > 0 ? __float2int_ru(q) : __float2int_rd(q);
                            g = min(max(g, lo - (st < 0)), hi + (st > 0));
                            const unsigned span = (unsigned)(hi - lo);
                            while ((unsigned)(g - st - lo) <= span && !(sg * ((ak * (x0 + (float)(g - st)) + tk) + ck) < 0.0f)) g -= st;
                            while ((unsigned)(g - lo) <= span && (sg * ((ak * (x0 + (float)g) + tk) + ck) < 0.0f)) g += st;
                            if (st > 0) lo = g; else hi = g;
                        }
                    } else if (sg * ((ak * x0 + tk) + ck) < 0.0f) { // constant along the row (ak == 0 or NaN): all or nothing
                        hi = lo - 1;
                    }
                }
            }
        }
        // push the spans, at most 4 cells per lane per pass, and drain full warps of candidates.  The stack is an
        // unordered pool, so the pass's cells go in "column-major": all first cells, then all second cells, ... -- one
        // ballot + popc per column instead of a prefix scan and a divergent store loop.
        int remaining = hi >= lo ? hi - lo + 1 : 0;
        const unsigned lt_mask = (1u << lane) - 1u;
        while (__any_sync(FULL, remaining > 0)) {
            const int n = min(remaining, 4);
            const unsigned head = ((unsigned)slot << 24) | ((unsigned)lrow << 12);
            int ptot = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned bj = __ballot_sync(FULL, n > j);
                if (n > j) my_ring[pending + ptot + __popc(bj & lt_mask)] = head | (unsigned)(lo + j);
                ptot += __popc(bj);
            }
            lo += n; remaining -= n; pending += ptot;
            __syncwarp();
            while (pending >= 32) {
                pending -= 32;
                cover_from_slot<SC>(my_slots, my_ring[pending + lane], key, W, H, a);
            }
            __syncwarp();
        }
    }
    if (lane < pending) cover_from_slot<SC>(my_slots, my_ring[lane], key, W, H, a);
}

template <int SHADER, bool SC>
__global__ void __launch_bounds__(RW * 32, RT_RASTER_MINB) raster_kernel(const DrawArgs a, WorkCtl *ctl, uint2 *items, Slot *bigslots, const unsigned capacity)
{
    __shared__ Slot slots[RW][32];
    __shared__ unsigned ring[RW][RING];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int W = a.width, H = a.height;
    Slot *my_slots = slots[wid];
    unsigned *my_ring = ring[wid];

    const long long t = ((long long)blockIdx.x * RW + wid) * 32 + lane;
    VO q[2][3];
    int nprim = 0;
    if (t < a.n_tris) nprim = assemble<SHADER>(a, t, q);

#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (k == 1 && !__any_sync(FULL, nprim > 1)) break; // second output triangles: near-plane only
        Slot mine;
        int ncells = 0;
        if (k < nprim) ncells = setup_prim<SHADER, SC>(a, q[k][0], q[k][1], q[k][2], (unsigned)(2 * t + k), mine);

        // large primitives -> global work queue (falls back to inline if the queue is full)
        const unsigned big = __ballot_sync(FULL, ncells > SMALL_MAX);
        if (big) {
            // work items count row-quads: ceil(nx/4) * ny quads, QUADS per item
            const int nchunks = ncells > SMALL_MAX ? ((((mine.nx_tle & 0xffff) + 3) >> 2) * mine.off + QUADS - 1) / QUADS : 0;
            int incl = nchunks;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                int v = __shfl_up_sync(FULL, incl, d);
                if (lane >= d) incl += v;
            }
            const int tot = __shfl_sync(FULL, incl, 31);
            // n_items only grows (no give-back): every slot below min(n_items, capacity) is owned by exactly one warp.  A warp whose
            // range does not fit rasterizes its large primitives inline and fills what it reserved below the capacity with null
            // items (coverage_kernel skips them), so the consumer never reads an unwritten slot.
            unsigned base = capacity;
            if (lane == 0 && *(volatile unsigned *)&ctl->n_items < capacity) base = atomicAdd(&ctl->n_items, (unsigned)tot);
            base = __shfl_sync(FULL, base, 0);
            if (base + (unsigned)tot <= capacity) {
                unsigned w = base + (unsigned)(incl - nchunks);
                if (nchunks) {
                    bigslots[w] = mine; // setup travels with the first item: coverage_kernel recomputes nothing
                    for (int j = 0; j < nchunks; ++j) items[w + j] = make_uint2(w, (unsigned)j);
                    ncells = 0; // handed over
                }
            } else {
                for (unsigned w = base + (unsigned)lane; w < capacity; w += 32u) items[w] = make_uint2(0xffffffffu, 0u);
                if (lane == 0) atomicAdd(&ctl->overflowed, 1u);
            }
        }

        // ---- inline coverage by row spans ---------------------------------------------------------------------
        // The rows of the warp's staged primitives are laid end to end and taken 32 at a time, one row per lane.
        // A lane finds the exact interval of its row that the division-free sign test accepts (the float edge
        // functions are monotone in px, so per edge the accepted cells are a half-line: an analytic guess is
        // corrected by evaluating the real predicate), pushes those cells on the warp's candidate stack, and
        // full warps of candidates go through the exact test + depth atomics.
        const int ny = ncells > 0 ? mine.off : 0; // setup_prim left ny in .off
        const unsigned staged = __ballot_sync(FULL, ny > 0);
        if (staged == 0) continue;
        int incl = ny; // inclusive warp scan of row counts
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            int v = __shfl_up_sync(FULL, incl, d);
            if (lane >= d) incl += v;
        }
        const int off = incl - ny, total = __shfl_sync(FULL, incl, 31);
        __syncwarp(); // previous pass finished reading the slots
        if (ny > 0) {
            mine.off = off;
            my_slots[__popc(staged & lt_mask)] = mine;
        }
        __syncwarp();

        cover_rows<SC>(my_slots, my_ring, ny, off, total, lane, a.key, W, H, a);
    }
}

// ---- kernel 1b: coverage of the queued large primitives ---------------------------------------------
// Persistent grid; each warp takes whole work items round-robin.  An item is 32 row-quads of one primitive: a
// lane owns 4 horizontally adjacent cells, so the row term b*py of each edge function is shared (bit-identical:
// the reference evaluates a*px + b*py + c as (a*px + b*py) + c, and fl(b*py) does not depend on px).  Lanes
// sign-test their 4 cells; the survivors of the whole warp are then dealt out one per lane (ballot + find-nth-set)
// so the exact test -- 3 IEEE divisions, depth, 64-bit atomicMin -- always runs on full warps.
template <int SHADER, bool SC>
__global__ void __launch_bounds__(128) coverage_kernel(const DrawArgs a, const WorkCtl *ctl, const uint2 *items, const Slot *bigslots, const unsigned capacity)
{
    const unsigned FULL = 0xffffffffu;
    const unsigned n_raw = ctl->n_items, n_items = n_raw < capacity ? n_raw : capacity; // the producers' counter may run past the queue's capacity
    const int lane = threadIdx.x & 31;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const int W = a.width, H = a.height;
    for (unsigned i = warp; i < n_items; i += n_warps) {
        const uint2 it = items[i];
        if (it.x == 0xffffffffu) continue; // reserved by a warp whose range did not fit
        const float4 *sp = reinterpret_cast<const float4 *>(bigslots + it.x);
        const float4 s0 = __ldcg(sp), s1 = __ldcg(sp + 1), s2 = __ldcg(sp + 2), s3 = __ldcg(sp + 3), s4 = __ldcg(sp + 4), s5 = __ldcg(sp + 5);
        Edges e;
        e.a1 = s0.x; e.b1 = s0.y; e.c1 = s0.z; e.a2 = s0.w; e.b2 = s1.x; e.c2 = s1.y;
        e.a3 = s1.z; e.b3 = s1.w; e.c3 = s2.x;
        const int ny = __float_as_int(s2.z), sxy = __float_as_int(s2.w), nx_tle = __float_as_int(s3.x);
        e.tle = (unsigned)(nx_tle >> 16) & 7u;
        const int nx = nx_tle & 0xffff, startx = sxy & 0xffff, starty = sxy >> 16;
        const int nxq = (nx + 3) >> 2, nquads = nxq * ny;
        const unsigned prim = __float_as_uint(s3.y);

        const int q = (int)it.y * QUADS + lane;
        unsigned cand = 0; // bit c: cell (col0 + c, row) survives the sign test
        int col0 = 0, row = 0;
        if (q < nquads) {
            const int r = (int)(((float)q + 0.5f) * __frcp_rn((float)nxq)); // exact: q, nxq < 4096
            col0 = startx + 4 * (q - r * nxq);
            row = starty + r;
            const float py = (float)row + 0.5f;
            const float t1 = e.b1 * py, t2 = e.b2 * py, t3 = e.b3 * py;
            int ncol = min(4, startx + nx - col0);
            if (SC && (nx_tle & (1 << 21)) && !rt_owns_any_row(a.own, row - 1, row + 1)) ncol = 0; // see setup_prim: "local"
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float px = (float)(col0 + c) + 0.5f;
                const float d1 = e.a1 * px + t1 + e.c1, d2 = e.a2 * px + t2 + e.c2, d3 = e.a3 * px + t3 + e.c3;
                const float s = d1 + d2 + d3;
                if (c < ncol && !(d1 * s < 0.0f || d2 * s < 0.0f || d3 * s < 0.0f)) cand |= 1u << c;
            }
        }
        // deal the warp's candidates out, 32 per round
        const unsigned m0 = __ballot_sync(FULL, cand & 1u), m1 = __ballot_sync(FULL, cand & 2u), m2 = __ballot_sync(FULL, cand & 4u),
                       m3 = __ballot_sync(FULL, cand & 8u);
        const int c0 = __popc(m0), c1 = c0 + __popc(m1), c2 = c1 + __popc(m2), total = c2 + __popc(m3);
        for (int base = 0; base < total; base += 32) {
            const int k = base + lane;
            const bool act = k < total;
            int c = 0, src = 0;
            if (act) {
                c = k < c0 ? 0 : k < c1 ? 1 : k < c2 ? 2 : 3;
                const unsigned m = c == 0 ? m0 : c == 1 ? m1 : c == 2 ? m2 : m3;
                const int before = c == 0 ? 0 : c == 1 ? c0 : c == 2 ? c1 : c2;
                src = (int)__fns(m, 0, k - before + 1); // lane that owns this candidate
            }
            const int scol = __shfl_sync(FULL, col0, src), srow = __shfl_sync(FULL, row, src);
            if (act) cover_cell<SC>(e, scol + c, srow, s3.z, s3.w, s4.x, s4.y, s4.z, s4.w, s5.x, s5.y, s5.z, prim, a.key, W, H, a);
        }
    }
}

// ---- kernel 2: resolve + fragment shader ----------------------------------------------------------

struct ResolveArgs {
    WorkCtl *ctl;
    unsigned long long *key;
    const float4 *rec;
    uint32_t *bgra;
    int width, height;
    cudaTextureObject_t tex;
    int tex_w, tex_h;
    int clear;         // 1: a clear(render_target) is folded into this draw ...
    uint32_t clear_px; // ... pixels nobody wins get this BGRA8 value
    RtOwner own;       // pixels this draw may write (resolve, clear colour, key re-arm) ...
    int scissor;       // ... 0: the whole frame
    int bx0, by0;      // first 32x32 block of the launch grid (the grid covers the owned rect only)
};

struct PrimRec { float4 h1, h2, h3; };

// Slow path (practically never taken): the winning fragment was produced by a loop cell other than the
// pixel itself, because (int) of the interpolated position landed elsewhere.  Search the bbox in row-major
// order for the first cell of this primitive that maps to (x,y) with these depth bits.
__device__ __noinline__ bool find_source_cell(const PrimRec &p, const Edges &e, int x, int y, uint32_t zbits, int W, int H,
                                               int *col_out, int *row_out)
{
    BBox bb = bbox_setup(p.h1.x, p.h1.y, p.h2.x, p.h2.y, p.h3.x, p.h3.y, W, H);
    for (int r = 0; r < bb.ny; ++r)
        for (int c = 0; c < bb.nx; ++c) {
            int col = bb.startx + c, row = bb.starty + r;
            Cell cl = cell_eval(e, col, row);
            if (!cl.inside) continue;
            float hx = blend3(p.h1.x, p.h2.x, p.h3.x, cl.al1, cl.al2, cl.al3);
            float hy = blend3(p.h1.y, p.h2.y, p.h3.y, cl.al1, cl.al2, cl.al3);
            float hz = blend3(p.h1.z, p.h2.z, p.h3.z, cl.al1, cl.al2, cl.al3);
            if (hz < 0 || (int)hx != x || (int)hy != y || __float_as_uint(hz) != zbits) continue;
            *col_out = col; *row_out = row;
            return true;
        }
    return false;
}

// Shade one pixel whose key names a winner of this draw.
template <int SHADER>
__device__ __forceinline__ void resolve_pixel(const ResolveArgs &a, int x, int y, unsigned long long k64)
{
    const size_t p = (size_t)y * a.width + x;
    const unsigned prim = (unsigned)k64;
    const uint32_t zbits = (uint32_t)(k64 >> 32);
    const float4 *r = a.rec + (size_t)prim * RecLayout<SHADER>::F4;
    PrimRec pr;
    pr.h1 = __ldg(r); pr.h2 = __ldg(r + 1); pr.h3 = __ldg(r + 2);
    const float4 at = __ldg(r + 3);
    Edges e = edge_setup(pr.h1.x, pr.h1.y, pr.h2.x, pr.h2.y, pr.h3.x, pr.h3.y);

    int col = x, row = y;
    Cell cl = cell_eval(e, col, row);
    {
        float hx = blend3(pr.h1.x, pr.h2.x, pr.h3.x, cl.al1, cl.al2, cl.al3);
        float hy = blend3(pr.h1.y, pr.h2.y, pr.h3.y, cl.al1, cl.al2, cl.al3);
        float hz = blend3(pr.h1.z, pr.h2.z, pr.h3.z, cl.al1, cl.al2, cl.al3);
        if (!(cl.inside && (int)hx == x && (int)hy == y && __float_as_uint(hz) == zbits)) {
            if (find_source_cell(pr, e, x, y, zbits, a.width, a.height, &col, &row)) cl = cell_eval(e, col, row);
        }
    }
    // _raster.py:313-318 perspective-correct weights, interpolate3 with (beta2, beta3)
    float q1 = cl.al1 / pr.h1.w, q2 = cl.al2 / pr.h2.w, q3 = cl.al3 / pr.h3.w;
    float qs = q1 + q2 + q3;
    float beta2 = q2 / qs, beta3 = q3 / qs;
    float w1 = 1.0f - beta2 - beta3;
    float4 color;
    if (SHADER == RT_SHADER_LESSON08) { // lesson08:58-62
        float c = blend3(at.x, at.y, at.z, w1, beta2, beta3);
        color = make_float4(c, c, c, 1.0f);
    } else {                            // lesson09:90-95
        const float4 uv12 = __ldg(r + 4), uv3 = __ldg(r + 5);
        float L = blend3(at.x, at.y, at.z, w1, beta2, beta3);
        float cx = blend3(uv12.x, uv12.z, uv3.x, w1, beta2, beta3);
        float cy = blend3(uv12.y, uv12.w, uv3.y, w1, beta2, beta3);
        float4 tx = rt_sample2d(a.tex, a.tex_w, a.tex_h, cx, cy);
        color = make_float4(tx.x * L, tx.y * L, tx.z * L, 1.0f);
    }
    const float z = __uint_as_float(zbits);
    if (!(z <= 0)) a.bgra[p] = rt_pack_bgra(color.x, color.y, color.z, color.w); // FragmentProcess, :102
    else if (a.clear) a.bgra[p] = a.clear_px;
    a.key[p] = k64 | 0xFFFFFFFFull; // re-arm: later draws win depth ties, as in the reference's draw order
}

// 256 threads cover a 32x32 pixel block; each warp an 8x16 strip of which every lane owns 4 pixels (rows y, y+4,
// y+8, y+12).  The four key loads are issued before any is used: the kernel is bound by the latency of that one
// dependent load per pixel, so memory-level parallelism is what buys time here.
template <int SHADER>
__global__ void __launch_bounds__(256, RT_RESOLVE_MINB) resolve_kernel(const ResolveArgs a)
{
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) a.ctl->n_items = 0; // queue drained: re-arm for the next draw
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int x = (a.bx0 + blockIdx.x) * 32 + (wid & 3) * 8 + (lane & 7), y0 = (a.by0 + blockIdx.y) * 32 + (wid >> 2) * 16 + (lane >> 3);
    if (x >= a.width) return;
    if (a.scissor && (x < a.own.x0 || x > a.own.x1 || !rt_owns_any_row(a.own, y0, y0 + 12))) return;
    unsigned long long k[4];
    bool mine[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int y = y0 + 4 * i;
        mine[i] = y < a.height && (!a.scissor || rt_owns_row(a.own, y));
        k[i] = mine[i] ? __ldcs(a.key + (size_t)y * a.width + x) : ~0ull;
    }
#ifndef RT_RESOLVE_PREFETCH
#define RT_RESOLVE_PREFETCH 1
#endif
#if RT_RESOLVE_PREFETCH
    // start the four record fetches before the first pixel is shaded: the loop below is one dependent key -> record chain per
    // pixel.  Measured on B200 (cfg2, 8 frame streams): 49.8 us per frame without, 48.5 with prefetch.global.L1, 48.6 with .L2
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if ((unsigned)k[i] != RT_NO_PRIMITIVE) {
            const float4 *r = a.rec + (size_t)(unsigned)k[i] * RecLayout<SHADER>::F4;
            rt_prefetch<RT_RESOLVE_PREFETCH>(r);
        }
#endif
    // Measured on B200 and NOT adopted (dragon100k, 1080p, frame 96.0 us with this form): 1 or 2 pixels per thread and/or 4-6
    // resident blocks per SM via __launch_bounds__ (64 / 48 / 40 registers): 96.5-114 us -- the register caps spill, and
    // fewer pixels per thread lose the memory-level parallelism of the four up-front key loads.
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const unsigned long long k64 = i == 0 ? k[0] : i == 1 ? k[1] : i == 2 ? k[2] : k[3];
        const bool own_i = i == 0 ? mine[0] : i == 1 ? mine[1] : i == 2 ? mine[2] : mine[3];
        if ((unsigned)k64 != RT_NO_PRIMITIVE) resolve_pixel<SHADER>(a, x, y0 + 4 * i, k64);
        else if (a.clear && own_i) a.bgra[(size_t)(y0 + 4 * i) * a.width + x] = a.clear_px;
    }
}

// ---- Raster.draw_points (rendering/_raster.py:399-414) -------------------------------------------------------------
// One thread per point: vertex shader, PointAssembly's z<0 cull (:146), PointRaster's clip-space |x|,|y| <= w test
// (:221), Dehomogenize, DepthTest as a 64-bit atomicMin on (depth bits, point id); a 32 B record per point keeps the
// vertex output for the resolve pass.
template <int SHADER>
__global__ void __launch_bounds__(256) points_kernel(const DrawArgs a)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_tris) return; // n_tris carries the point count here
    const long long vi = a.idx ? (long long)a.idx[i] : i;
    VO v = vertex_shader<SHADER>(__ldg(a.pos + vi), __ldg(a.nrm + vi), a.g);
    if (v.z < 0) return;
    if (v.x < -v.w || v.x > v.w || v.y < -v.w || v.y > v.w) return;
    dehomogenize(v, a.half_w, a.half_h);
    a.rec[2 * i] = make_float4(v.x, v.y, v.z, v.w);
    a.rec[2 * i + 1] = make_float4(v.a0, v.a1, v.a2, 0.0f);
    if (v.z < 0) return;
    const int ix = (int)v.x, iy = (int)v.y;
    if (ix < 0 || ix >= a.width || iy < 0 || iy >= a.height) return;
    if (a.scissor && !rt_owns(a.own, ix, iy)) return;
    atomicMin(a.key + (size_t)iy * a.width + ix, ((unsigned long long)__float_as_uint(v.z) << 32) | (unsigned)i);
}

template <int SHADER>
__global__ void __launch_bounds__(256) resolve_points_kernel(const ResolveArgs a)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)a.width * a.height) return;
    if (a.scissor && !rt_owns(a.own, (int)(p % a.width), (int)(p / a.width))) return;
    const unsigned long long k64 = a.key[p];
    const unsigned prim = (unsigned)k64;
    if (prim == RT_NO_PRIMITIVE) {
        if (a.clear) a.bgra[p] = a.clear_px;
        return;
    }
    const float4 at = __ldg(a.rec + 2 * (size_t)prim + 1);
    float4 color;
    if (SHADER == RT_SHADER_LESSON08) {
        color = make_float4(at.x, at.x, at.x, 1.0f);
    } else {
        float4 tx = rt_sample2d(a.tex, a.tex_w, a.tex_h, at.y, at.z);
        color = make_float4(tx.x * at.x, tx.y * at.x, tx.z * at.x, 1.0f);
    }
    const float z = __uint_as_float((uint32_t)(k64 >> 32));
    if (!(z <= 0)) a.bgra[p] = rt_pack_bgra(color.x, color.y, color.z, color.w);
    else if (a.clear) a.bgra[p] = a.clear_px;
    a.key[p] = k64 | 0xFFFFFFFFull;
}

// ---- clears and depth views -----------------------------------------------------------------------

__global__ void fill_u64_kernel(ulonglong2 *dst, long long n2, unsigned long long v, unsigned long long *tail, int ntail)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const ulonglong2 vv = make_ulonglong2(v, v);
    for (; i < n2; i += stride) dst[i] = vv;
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = v;
}

__global__ void fill_u32_kernel(uint4 *dst, long long n4, uint32_t v, uint32_t *tail, int ntail)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const uint4 vv = make_uint4(v, v, v, v);
    for (; i < n4; i += stride) dst[i] = vv;
    if (blockIdx.x == 0 && threadIdx.x < ntail) tail[threadIdx.x] = v;
}

// clear(depth_buffer) under a scissor / stripe ownership: only owned pixels (one block per row of the owned rect)
__global__ void fill_u64_owned_kernel(unsigned long long *key, int width, const RtOwner own, unsigned long long v)
{
    const int y = own.y0 + (int)blockIdx.x;
    if (!rt_owns_row(own, y)) return;
    unsigned long long *row = key + (size_t)y * width;
    for (int x = own.x0 + (int)threadIdx.x; x <= own.x1; x += (int)blockDim.x) row[x] = v;
}

__global__ void read_depth_kernel(const unsigned long long *key, long long n, uint32_t *out)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (uint32_t)(key[i] >> 32);
}

__global__ void write_depth_kernel(unsigned long long *key, long long n, const uint32_t *in)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) key[i] = ((unsigned long long)in[i] << 32) | 0xFFFFFFFFull;
}

// host-side replica of rt_pack_bgra (sat, x255, round to nearest even)
uint32_t pack_bgra_host(const float rgba[4])
{
    uint32_t px = 0;
    const int order[4] = {2, 1, 0, 3};
    for (int i = 0; i < 4; ++i) {
        float v = rgba[order[i]] * 255.0f;
        if (!(v > 0.0f)) v = 0.0f;
        if (v > 255.0f) v = 255.0f;
        px |= (uint32_t)__builtin_nearbyintf(v) << (8 * i);
    }
    return px;
}

int fill_grid(long long n_vec)
{
    long long blocks = (n_vec + 255) / 256;
    long long cap = (long long)rt_sm_count() * 8;
    return (int)(blocks < 1 ? 1 : (blocks > cap ? cap : blocks));
}

constexpr size_t CTL_BYTES = 256;

template <int SHADER>
int launch_draw(const DrawArgs &da, ResolveArgs ra, void *scratch, long long scratch_bytes, cudaStream_t st)
{
    // scratch = [WorkCtl, 256 B][records: 2 per triangle][big-primitive slots: cap x 96 B][work items: cap x 8 B]
    char *base = (char *)scratch;
    WorkCtl *ctl = (WorkCtl *)base;
    const long long rec_bytes = 2 * da.n_tris * RecLayout<SHADER>::F4 * 16;
    long long cap = (scratch_bytes - (long long)CTL_BYTES - rec_bytes) / (long long)(sizeof(uint2) + sizeof(Slot));
    if (cap < 0) cap = 0;
    if (cap > 0x7fffffffll) cap = 0x7fffffffll;
    Slot *bigslots = (Slot *)(base + CTL_BYTES + rec_bytes);
    uint2 *items = (uint2 *)(base + CTL_BYTES + rec_bytes + cap * (long long)sizeof(Slot));
    ra.ctl = ctl;
    if (da.n_tris > 0) {
        // one warp per 32 triangles, RW warps per block; the hardware block scheduler does the load balancing
        long long blocks = (da.n_tris + RW * 32 - 1) / (RW * 32);
        if (da.scissor) raster_kernel<SHADER, true><<<(unsigned)blocks, RW * 32, 0, st>>>(da, ctl, items, bigslots, (unsigned)cap);
        else raster_kernel<SHADER, false><<<(unsigned)blocks, RW * 32, 0, st>>>(da, ctl, items, bigslots, (unsigned)cap);
        RT_CUDA(cudaGetLastError());
        if (cap > 0) {
            if (da.scissor) coverage_kernel<SHADER, true><<<rt_sm_count() * RT_COVERAGE_BLOCKS_PER_SM, 128, 0, st>>>(da, ctl, items, bigslots, (unsigned)cap);
            else coverage_kernel<SHADER, false><<<rt_sm_count() * RT_COVERAGE_BLOCKS_PER_SM, 128, 0, st>>>(da, ctl, items, bigslots, (unsigned)cap);
            RT_CUDA(cudaGetLastError());
        }
    }
    ra.bx0 = ra.own.x0 / 32; ra.by0 = ra.own.y0 / 32; // own is the whole frame without a scissor
    if (ra.own.x1 < ra.own.x0 || ra.own.y1 < ra.own.y0) { ra.bx0 = ra.by0 = 0; ra.own.x1 = ra.own.x0 = 0; ra.own.y1 = ra.own.y0 = 0; ra.own.mod = 2u; ra.own.rem = 1u; ra.own.rows = 1u << 30; } // empty: one block, owns nothing
    dim3 grid(ra.own.x1 / 32 - ra.bx0 + 1, ra.own.y1 / 32 - ra.by0 + 1), block(256);
    resolve_kernel<SHADER><<<grid, block, 0, st>>>(ra);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

} // namespace

extern "C" {

int rt_raster_clear_depth(void *d_key, int64_t n_pixels, uint32_t depth_bits, void *stream)
{
    RT_REQUIRE(d_key && n_pixels >= 0, "key buffer");
    RT_REQUIRE(((uintptr_t)d_key & 15) == 0, "key buffer must be 16-byte aligned");
    if (n_pixels == 0) return RT_OK;
    unsigned long long v = ((unsigned long long)depth_bits << 32) | 0xFFFFFFFFull;
    long long n2 = n_pixels / 2;
    fill_u64_kernel<<<fill_grid(n2), 256, 0, (cudaStream_t)stream>>>((ulonglong2 *)d_key, n2, v,
                                                                    (unsigned long long *)d_key + 2 * n2, (int)(n_pixels - 2 * n2));
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

} // extern "C"

// clear(depth_buffer) folded into a draw: the whole key buffer, or under a scissor only the pixels the draw owns
static int clear_depth_owned(void *d_key, int width, int height, uint32_t depth_bits, const RtOwner &own, int scissor, void *stream)
{
    if (!scissor) return rt_raster_clear_depth(d_key, (int64_t)width * height, depth_bits, stream);
    if (own.x1 < own.x0 || own.y1 < own.y0) return RT_OK;
    const unsigned long long v = ((unsigned long long)depth_bits << 32) | 0xFFFFFFFFull;
    fill_u64_owned_kernel<<<(unsigned)(own.y1 - own.y0 + 1), 256, 0, (cudaStream_t)stream>>>((unsigned long long *)d_key, width, own, v);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

extern "C" {

int rt_raster_clear_color(void *d_bgra, int64_t n_pixels, const float rgba[4], void *stream)
{
    RT_REQUIRE(d_bgra && rgba && n_pixels >= 0, "colour buffer");
    RT_REQUIRE(((uintptr_t)d_bgra & 15) == 0, "colour buffer must be 16-byte aligned");
    if (n_pixels == 0) return RT_OK;
    const uint32_t px = pack_bgra_host(rgba);
    long long n4 = n_pixels / 4;
    fill_u32_kernel<<<fill_grid(n4), 256, 0, (cudaStream_t)stream>>>((uint4 *)d_bgra, n4, px, (uint32_t *)d_bgra + 4 * n4,
                                                                    (int)(n_pixels - 4 * n4));
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int rt_raster_read_depth(const void *d_key, int64_t n_pixels, void *d_depth_u32, void *stream)
{
    RT_REQUIRE(d_key && d_depth_u32 && n_pixels >= 0, "buffers");
    if (n_pixels == 0) return RT_OK;
    read_depth_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const unsigned long long *)d_key, n_pixels,
                                                                                           (uint32_t *)d_depth_u32);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int rt_raster_write_depth(void *d_key, int64_t n_pixels, const void *d_depth_u32, void *stream)
{
    RT_REQUIRE(d_key && d_depth_u32 && n_pixels >= 0, "buffers");
    if (n_pixels == 0) return RT_OK;
    write_depth_kernel<<<(unsigned)((n_pixels + 255) / 256), 256, 0, (cudaStream_t)stream>>>((unsigned long long *)d_key, n_pixels,
                                                                                            (const uint32_t *)d_depth_u32);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int64_t rt_raster_scratch_bytes(int shader, int64_t n_triangles, int width, int height)
{
    const int64_t f4 = shader == RT_SHADER_LESSON08 ? 4 : 6;
    const int64_t rec = 2 * n_triangles * f4 * 16; // up to two primitives per input triangle
    // work items: one per 32 row-quads (<= 128 cells) of a large primitive, each with room for a 96 B setup slot.
    // The reference caps one pass at 32*W*H bbox cells (_raster.py:378); sized for a quarter of that plus one
    // partial item per primitive, bounded to 2^21 items (208 MB).  A full queue only means inline (slower) coverage.
    int64_t items = 8ll * width * height / (4 * QUADS) + (2 * n_triangles < (1ll << 20) ? 2 * n_triangles : (1ll << 20));
    if (items > (1ll << 21)) items = 1ll << 21;
    return (int64_t)CTL_BYTES + rec + items * (int64_t)(sizeof(uint2) + sizeof(Slot));
}

int rt_raster_draw_triangles(const void *d_pos4, const void *d_nrm4, const int32_t *d_indices, int64_t n_triangles, int shader,
                             const float *vs_globals, uint64_t tex_handle, int width, int height, void *d_key, void *d_scratch,
                             int64_t scratch_bytes, void *d_bgra, const float *clear_rgba, int clear_depth, uint32_t clear_depth_bits,
                             const int32_t *owner, void *stream)
{
    RT_REQUIRE(n_triangles >= 0 && n_triangles < (1ll << 31), "triangle count (primitive ids are 32-bit: 2*t+k)");
    RT_REQUIRE(n_triangles == 0 || (d_pos4 && d_nrm4), "vertex arrays");
    RT_REQUIRE(d_scratch && ((uintptr_t)d_scratch & 15) == 0, "scratch buffer (16-byte aligned)");
    RT_REQUIRE(scratch_bytes >= (int64_t)CTL_BYTES + 2 * n_triangles * (shader == RT_SHADER_LESSON08 ? 4 : 6) * 16,
               "scratch buffer too small: see rt_raster_scratch_bytes");
    RT_REQUIRE(vs_globals && d_key && d_bgra, "globals / targets");
    RT_REQUIRE(width > 0 && height > 0 && width <= 32768 && height <= 32768, "viewport");
    RT_REQUIRE(shader == RT_SHADER_LESSON08 || shader == RT_SHADER_LESSON09, "shader id");
    DrawArgs da;
    da.pos = (const float4 *)d_pos4; da.nrm = (const float4 *)d_nrm4; da.idx = d_indices; da.n_tris = n_triangles;
    for (int i = 0; i < 48; ++i) da.g[i] = vs_globals[i];
    da.width = width; da.height = height;
    da.half_w = (float)width * 0.5f; da.half_h = (float)height * 0.5f; // viewport_dim * 0.5f, _raster.py:129
    da.key = (unsigned long long *)d_key; da.rec = (float4 *)((char *)d_scratch + CTL_BYTES);
    da.scissor = rt_owner_parse(owner, width, height, &da.own);
    RT_REQUIRE(da.scissor >= 0, "owner: {x0, y0, x1, y1, stripe rows >= 1, stripe mod >= 1, 0 <= stripe rem < mod}");
    ResolveArgs ra;
    ra.ctl = nullptr;
    ra.key = da.key; ra.rec = da.rec; ra.bgra = (uint32_t *)d_bgra; ra.width = width; ra.height = height;
    ra.tex = 0; ra.tex_w = 0; ra.tex_h = 0;
    ra.clear = clear_rgba ? 1 : 0;
    ra.clear_px = clear_rgba ? pack_bgra_host(clear_rgba) : 0u;
    ra.own = da.own; ra.scissor = da.scissor; ra.bx0 = ra.by0 = 0;
    if (clear_depth) { // a pending clear(depth_buffer, v) folded into this call: must precede the coverage atomics
        int rc = clear_depth_owned(d_key, width, height, clear_depth_bits, da.own, da.scissor, stream);
        if (rc != RT_OK) return rc;
    }
    if (shader == RT_SHADER_LESSON09) {
        RT_REQUIRE(tex_handle != 0, "lesson09 shader needs a texture handle");
        const rt_texture *t = (const rt_texture *)(uintptr_t)tex_handle;
        ra.tex = t->obj; ra.tex_w = t->w; ra.tex_h = t->h;
        return launch_draw<RT_SHADER_LESSON09>(da, ra, d_scratch, scratch_bytes, (cudaStream_t)stream);
    }
    return launch_draw<RT_SHADER_LESSON08>(da, ra, d_scratch, scratch_bytes, (cudaStream_t)stream);
}

int64_t rt_raster_points_scratch_bytes(int64_t n_points) { return (int64_t)CTL_BYTES + 32 * n_points; }

int rt_raster_draw_points(const void *d_pos4, const void *d_nrm4, const int32_t *d_indices, int64_t n_points, int shader,
                          const float *vs_globals, uint64_t tex_handle, int width, int height, void *d_key, void *d_scratch,
                          int64_t scratch_bytes, void *d_bgra, const float *clear_rgba, int clear_depth, uint32_t clear_depth_bits,
                          const int32_t *owner, void *stream)
{
    RT_REQUIRE(n_points >= 0 && n_points < 0xFFFFFFFFll, "point count (ids are 32-bit)");
    RT_REQUIRE(n_points == 0 || (d_pos4 && d_nrm4), "vertex arrays");
    RT_REQUIRE(vs_globals && d_key && d_bgra && d_scratch, "globals / targets / scratch");
    RT_REQUIRE(scratch_bytes >= rt_raster_points_scratch_bytes(n_points), "scratch buffer too small: see rt_raster_points_scratch_bytes");
    RT_REQUIRE(width > 0 && height > 0 && width <= 32768 && height <= 32768, "viewport");
    RT_REQUIRE(shader == RT_SHADER_LESSON08 || shader == RT_SHADER_LESSON09, "shader id");
    cudaStream_t st = (cudaStream_t)stream;
    DrawArgs da;
    da.pos = (const float4 *)d_pos4; da.nrm = (const float4 *)d_nrm4; da.idx = d_indices; da.n_tris = n_points;
    for (int i = 0; i < 48; ++i) da.g[i] = vs_globals[i];
    da.width = width; da.height = height;
    da.half_w = (float)width * 0.5f; da.half_h = (float)height * 0.5f;
    da.key = (unsigned long long *)d_key; da.rec = (float4 *)((char *)d_scratch + CTL_BYTES);
    da.scissor = rt_owner_parse(owner, width, height, &da.own);
    RT_REQUIRE(da.scissor >= 0, "owner: {x0, y0, x1, y1, stripe rows >= 1, stripe mod >= 1, 0 <= stripe rem < mod}");
    ResolveArgs ra;
    ra.own = da.own; ra.scissor = da.scissor; ra.bx0 = ra.by0 = 0;
    ra.ctl = (WorkCtl *)d_scratch; ra.key = da.key; ra.rec = da.rec; ra.bgra = (uint32_t *)d_bgra; ra.width = width; ra.height = height;
    ra.tex = 0; ra.tex_w = 0; ra.tex_h = 0;
    ra.clear = clear_rgba ? 1 : 0;
    ra.clear_px = clear_rgba ? pack_bgra_host(clear_rgba) : 0u;
    if (clear_depth) {
        int rc = clear_depth_owned(d_key, width, height, clear_depth_bits, da.own, da.scissor, stream);
        if (rc != RT_OK) return rc;
    }
    const unsigned pblocks = (unsigned)((n_points + 255) / 256), rblocks = (unsigned)(((long long)width * height + 255) / 256);
    if (shader == RT_SHADER_LESSON09) {
        RT_REQUIRE(tex_handle != 0, "lesson09 shader needs a texture handle");
        const rt_texture *t = (const rt_texture *)(uintptr_t)tex_handle;
        ra.tex = t->obj; ra.tex_w = t->w; ra.tex_h = t->h;
        if (n_points) points_kernel<RT_SHADER_LESSON09><<<pblocks, 256, 0, st>>>(da);
        resolve_points_kernel<RT_SHADER_LESSON09><<<rblocks, 256, 0, st>>>(ra);
    } else {
        if (n_points) points_kernel<RT_SHADER_LESSON08><<<pblocks, 256, 0, st>>>(da);
        resolve_points_kernel<RT_SHADER_LESSON08><<<rblocks, 256, 0, st>>>(ra);
    }
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

int rt_texture_create(const void *d_texels, int width, int height, uint64_t *out_handle)
{
    RT_REQUIRE(d_texels && out_handle && width > 0 && height > 0, "texture arguments");
    RT_REQUIRE(((uintptr_t)d_texels & 511) == 0, "texel pointer must be 512-byte aligned");
    RT_REQUIRE((long long)width * height <= (1ll << 27), "linear texture limited to 2^27 texels");
    cudaResourceDesc rd = {};
    rd.resType = cudaResourceTypeLinear;
    rd.res.linear.devPtr = const_cast<void *>(d_texels);
    rd.res.linear.desc = cudaCreateChannelDesc<float4>();
    rd.res.linear.sizeInBytes = (size_t)width * height * 16;
    cudaTextureDesc td = {};
    td.filterMode = cudaFilterModePoint;
    td.readMode = cudaReadModeElementType;
    td.normalizedCoords = 0;
    rt_texture *t = new rt_texture{0, width, height};
    cudaError_t e = cudaCreateTextureObject(&t->obj, &rd, &td, nullptr);
    if (e != cudaSuccess) {
        delete t;
        rt_set_error("cudaCreateTextureObject: %s", cudaGetErrorString(e));
        return RT_ERR_CUDA;
    }
    *out_handle = (uint64_t)(uintptr_t)t;
    return RT_OK;
}

int rt_texture_destroy(uint64_t handle)
{
    if (!handle) return RT_OK;
    rt_texture *t = (rt_texture *)(uintptr_t)handle;
    cudaDestroyTextureObject(t->obj);
    delete t;
    return RT_OK;
}

} // extern "C"
