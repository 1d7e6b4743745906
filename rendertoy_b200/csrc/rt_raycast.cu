// rt_raycast.cu -- Raycaster.ray_cast for sm_100a (rendering/_raycaster.py:35-36 is `pass` in the reference;
// semantics are those of oracle/raycast_oracle.c: float32 Moller-Trumbore without FMA, closest hit = min over
// (bits(t) << 32 | triangle id)).
//
// Primary rays (rt_raycast_primary), three launches per frame on the caller's stream:
//   project_kernel     every inner node -> its two children's screen rectangles + minimum depth for THIS camera
//   view_refit_kernel  a few in-place tightening iterations (a child's rectangle := union of its own children's)
//   raycast_kernel     one warp per 8x4-pixel tile, scheduled by the hardware: the 32 rays of a tile walk the screen-space
//                      nodes TOGETHER (one shared-memory stack per warp, broadcast 128-bit loads, votes), leaves run the
//                      exact Moller-Trumbore, the hit is shaded (Lambert / texture) straight into the BGRA8 frame -- per
//                      ray only 4 B (+16 B if hits are requested) leave the SM.  `stripes` restricts it to a rank's rows.
// Arbitrary rays (rt_raycast_rays) and scenes above 2^18 triangles walk the 3-D nodes per lane with a local-memory stack.
//
// Bound: instruction issue (0.82 of one instruction per scheduler and cycle with overlapped frames); the BVH of a
// 100k-triangle mesh is ~11 MB, L2-resident, HBM sees only the frame.  No contraction anywhere, so no tensor cores.
#include "rt_common.cuh"
#include "rt_bvh.cuh"

namespace {

#ifndef RT_RAYCAST_TB
#define RT_RAYCAST_TB 128
#endif
#ifndef RT_RAYCAST_MINB
#define RT_RAYCAST_MINB 12 // 40 registers: 48 of 64 warps resident (2 % over the unconstrained 46-register build)
#endif
constexpr int TB = RT_RAYCAST_TB;   // threads per block
constexpr int TILES_BX = TB >= 64 ? 2 : 1, TILES_BY = TB / 32 / TILES_BX; // tiles of one block
constexpr int STACK = 64; // Karras tree depth <= 64 (32 key bits + index tiebreak), one pending sibling per level
// The per-lane stack lives in local memory (L1-resident, only the touched depth is ever cached).  Measured on B200
// against a [depth][thread] shared-memory stack: 332 vs 374 us per 4K frame -- the 32 KB of shared memory per CTA
// cost more occupancy than the L1 round trips do.

// Primary rays share one origin, so a box test does not need the ray at all: with (a, b, c) the coordinates of a point in
// the camera's (U, V, W) basis, the ray through screen point (sx, sy) passes a point iff a / c = sx and b / c = sy, and
// its parameter there is t = c.  Once per frame project_kernel turns every inner node into the two children's screen
// rectangles (bounding the eight projected corners, padded by the direction rounding, see project_kernel) and minimum c;
// a node visit is then 3 loads + 10 compares instead of 4 loads + 12 FMA + 20 min/max.  The rectangle of a projected box
// is looser than the box (more visits), the exact Moller-Trumbore test at the leaves is unchanged, so are the hits.
struct __align__(16) ViewNode {
    float4 r0, r1; // child 0 / child 1: (sx_min, sx_max, sy_min, sy_max)
    float4 zc;     // (c_min child 0, c_min child 1, child0 bits, child1 bits)
};

struct ProjectArgs {
    const RtBvhNode *nodes;
    ViewNode *vnodes;
    long long n_inner;
    const RtBvhTri *tris;
    double minv[9]; // rows of [U V W]^-1
    double o[3];
    float pad_s;    // rectangle padding per unit of (1 + |s|)
    float rowsum;   // largest absolute row sum of [U V W]^-1: how far a unit displacement in space can move (a, b, c)
};

__device__ __forceinline__ void project_child(const ProjectArgs &p, float lox, float hix, float loy, float hiy, float loz, float hiz, float4 &rect,
                                              float &zmin)
{
    if (lox > hix) { // the empty box of a single-triangle scene
        rect = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
        zmin = INFINITY;
        return;
    }
    // (a, b, c) is affine in the corner: base + ix * ex + iy * ey + iz * ez
    const double qx = (double)lox - p.o[0], qy = (double)loy - p.o[1], qz = (double)loz - p.o[2];
    const double wx = (double)hix - (double)lox, wy = (double)hiy - (double)loy, wz = (double)hiz - (double)loz;
    double base[3], ex[3], ey[3], ez[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        base[r] = p.minv[3 * r] * qx + p.minv[3 * r + 1] * qy + p.minv[3 * r + 2] * qz;
        ex[r] = p.minv[3 * r] * wx; ey[r] = p.minv[3 * r + 1] * wy; ez[r] = p.minv[3 * r + 2] * wz;
    }
    float smin = INFINITY, smax = -INFINITY, tmin = INFINITY, tmax = -INFINITY, cmin = INFINITY, qmax = 0.0f;
    bool bad = false;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double a = base[0], b = base[1], c = base[2];
        if (k & 1) { a += ex[0]; b += ex[1]; c += ex[2]; }
        if (k & 2) { a += ey[0]; b += ey[1]; c += ey[2]; }
        if (k & 4) { a += ez[0]; b += ez[1]; c += ez[2]; }
        const float af = (float)a, bf = (float)b, cf = (float)c;
        const float s = af / cf, t = bf / cf; // float division: 2^-23 relative, far inside the padding
        smin = fminf(smin, s); smax = fmaxf(smax, s); tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
        cmin = fminf(cmin, cf);
        qmax = fmaxf(qmax, fmaxf(fabsf(af), fabsf(bf)));
        bad = bad || !(fabsf(s) < INFINITY) || !(fabsf(t) < INFINITY) || !(fabsf(cf) < INFINITY);
    }
    // a corner at or behind the eye plane (or non-finite data): every ray may pass -- no rectangle, no depth bound
    if (bad || !(cmin > 1e-6f * qmax) || !(cmin > 0.0f)) {
        rect = make_float4(-INFINITY, INFINITY, -INFINITY, INFINITY);
        zmin = 0.0f;
        return;
    }
    rect.x = smin - p.pad_s * (1.0f + fabsf(smin)); rect.y = smax + p.pad_s * (1.0f + fabsf(smax));
    rect.z = tmin - p.pad_s * (1.0f + fabsf(tmin)); rect.w = tmax + p.pad_s * (1.0f + fabsf(tmax));
    zmin = cmin * (1.0f - 64.0f * p.pad_s);
}

// A leaf's rectangle from the triangle itself: the box of a slanted triangle is mostly empty, and its projection more so.
// The triangle the exact test sees is v0 + u e1 + v e2; what it can report as a hit lies within the same 3-D tolerance
// the leaf boxes are padded with (read back here as box - vertex extent, doubled), which at depth c is a screen
// displacement of at most r rowsum (1 + |s|) / (c - r rowsum).  The result is intersected with the box's rectangle.
__device__ __forceinline__ void project_leaf(const ProjectArgs &p, int slot, float lox, float loy, float loz, float4 &rect, float &zmin)
{
    const float4 *tp = reinterpret_cast<const float4 *>(p.tris + slot);
    const float4 v0 = __ldg(tp), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
    const double P[3][3] = {{(double)v0.x, (double)v0.y, (double)v0.z},
                            {(double)v0.x + (double)e1.x, (double)v0.y + (double)e1.y, (double)v0.z + (double)e1.z},
                            {(double)v0.x + (double)e2.x, (double)v0.y + (double)e2.y, (double)v0.z + (double)e2.z}};
    const double vminx = fmin(P[0][0], fmin(P[1][0], P[2][0])), vminy = fmin(P[0][1], fmin(P[1][1], P[2][1])),
                 vminz = fmin(P[0][2], fmin(P[1][2], P[2][2]));
    const float r3 = 2.0f * (float)fmax(vminx - (double)lox, fmax(vminy - (double)loy, vminz - (double)loz)); // ~2 x the box padding
    float smin = INFINITY, smax = -INFINITY, tmin = INFINITY, tmax = -INFINITY, cmin = INFINITY;
    bool bad = !(r3 >= 0.0f) || !(r3 < INFINITY);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double qx = P[k][0] - p.o[0], qy = P[k][1] - p.o[1], qz = P[k][2] - p.o[2];
        const float af = (float)(p.minv[0] * qx + p.minv[1] * qy + p.minv[2] * qz), bf = (float)(p.minv[3] * qx + p.minv[4] * qy + p.minv[5] * qz),
                    cf = (float)(p.minv[6] * qx + p.minv[7] * qy + p.minv[8] * qz);
        const float s = af / cf, t = bf / cf;
        smin = fminf(smin, s); smax = fmaxf(smax, s); tmin = fminf(tmin, t); tmax = fmaxf(tmax, t);
        cmin = fminf(cmin, cf);
        bad = bad || !(fabsf(s) < INFINITY) || !(fabsf(t) < INFINITY) || !(fabsf(cf) < INFINITY);
    }
    const float reach = r3 * p.rowsum; // how far the tolerance moves (a, b, c)
    if (bad || !(cmin > 4.0f * reach) || !(cmin > 0.0f)) return; // keep the box's rectangle
    const float k = reach / (cmin - reach) * 1.0001f + p.pad_s;
    rect.x = fmaxf(rect.x, smin - k * (1.0f + fabsf(smin))); rect.y = fminf(rect.y, smax + k * (1.0f + fabsf(smax)));
    rect.z = fmaxf(rect.z, tmin - k * (1.0f + fabsf(tmin))); rect.w = fminf(rect.w, tmax + k * (1.0f + fabsf(tmax)));
    zmin = fmaxf(zmin, (cmin - reach) * (1.0f - 64.0f * p.pad_s));
}

// One thread per inner node.  Padding: the traced direction is fl(fl(U sx + V sy) + W), off the exact U sx + V sy + W by
// at most eps_d = 2^-22 (|U| |sx| + |V| |sy| + |W|) per component; through [U V W]^-1 that moves the ray's screen point by
// <= rowsum |Minv| eps_d (1 + |s|) -- the host passes pad_s = 8 x that bound (about 3 % of a 4K pixel for the lesson
// cameras).  The 3-D boxes already carry the 2^-17-extent padding that covers Moller-Trumbore's own rounding.
__global__ void __launch_bounds__(128) project_kernel(const ProjectArgs p)
{
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= p.n_inner) return;
    const float4 *np = reinterpret_cast<const float4 *>(p.nodes + i);
    const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
    const int4 n3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
    ViewNode v;
    float z0, z1;
    project_child(p, n0.x, n0.y, n0.z, n0.w, n2.x, n2.y, v.r0, z0);
    project_child(p, n1.x, n1.y, n1.z, n1.w, n2.z, n2.w, v.r1, z1);
    if (p.tris) {
        if (n3.x < 0) project_leaf(p, ~n3.x, n0.x, n0.z, n2.x, v.r0, z0);
        if (n3.y < 0 && n1.x <= n1.y) project_leaf(p, ~n3.y, n1.x, n1.z, n2.z, v.r1, z1);
    }
    v.zc = make_float4(z0, z1, __int_as_float(n3.x), __int_as_float(n3.y));
    p.vnodes[i] = v;
}

// project_kernel gives an inner child the rectangle of its projected 3-D box.  The union of that child's own two rectangles --
// which end, at the leaves, in triangle-tight rectangles -- is never larger and for slanted geometry much smaller, likewise
// the nearer of their depth bounds.  view_refit_kernel tightens every node from its children's CURRENT values, in place, `iters`
// times inside ONE launch (chaotic relaxation).  That is safe without any ordering or atomicity: a rectangle component only
// ever moves inwards and a depth bound only up, each stays a valid bound at every moment, so whatever mix of old and new
// values a thread reads (128-bit loads through L2, never L1) gives a valid, if looser, result.  A 100k-triangle tree is a
// single wave of threads, so an iteration costs one L2 round trip instead of a launch: tightness climbs at least one level
// per iteration (PLOC hands node ids out downwards, so with the reversed thread -> node mapping the blocks scheduled first
// hold the deepest nodes and it usually climbs several).  Hits cannot change: every leaf rectangle is untouched and every
// ancestor keeps containing it (tests: bit-identical hit records for 1..16 iterations, CPU emulator and B200).
// Measured on B200 (cfg4 frame / lesson08 camera at 4K, one frame alone): as separate launches 0 / 1 / 2 / 4 / 8 passes =
// 170 / 167 / 165 / 172 / 197 us and 408 / 373 / 358 / 347 / 358 us (profiles/r02a_ab_experimental.txt); in-kernel iterations: DESIGN.md 4.5.
__global__ void __launch_bounds__(128) view_refit_kernel(ViewNode *v, long long n_inner, int iters)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_inner) return;
    float4 *me = reinterpret_cast<float4 *>(v + (n_inner - 1 - t));
    float4 r[2] = {__ldcg(me), __ldcg(me + 1)};
    float4 zc = __ldcg(me + 2);
    const int child[2] = {__float_as_int(zc.z), __float_as_int(zc.w)};
    if (child[0] < 0 && child[1] < 0) return; // two leaves: already the triangles' rectangles
    float z[2] = {zc.x, zc.y};
    for (int it = 0; it < iters; ++it) {
        bool changed = false;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (child[c] < 0) continue; // a leaf: already the triangle's rectangle
            const float4 *cp = reinterpret_cast<const float4 *>(v + child[c]);
            const float4 a = __ldcg(cp), b = __ldcg(cp + 1), cz = __ldcg(cp + 2);
            // fminf / fmaxf drop NaN; an empty rectangle is (inf, -inf, inf, -inf) and an unbounded one (-inf, inf, -inf, inf)
            const float4 o = r[c];
            const float oz = z[c];
            r[c].x = fmaxf(o.x, fminf(a.x, b.x)); r[c].y = fminf(o.y, fmaxf(a.y, b.y));
            r[c].z = fmaxf(o.z, fminf(a.z, b.z)); r[c].w = fminf(o.w, fmaxf(a.w, b.w));
            z[c] = fmaxf(oz, fminf(cz.x, cz.y));
            changed = changed || r[c].x != o.x || r[c].y != o.y || r[c].z != o.z || r[c].w != o.w || z[c] != oz;
        }
        if (changed) {
            __stcg(me, r[0]); __stcg(me + 1, r[1]);
            __stcg(me + 2, make_float4(z[0], z[1], zc.z, zc.w));
        }
    }
}

int g_view_refit_passes = 4; // rt_raycast_set_view_refit; 4 iterations measured best on B200 for the cfg4 orbit (76.2 against 70.9 Grays/s without), 8 for the frame-filling camera

struct TraceArgs {
    const RtBvhNode *nodes;
    const RtBvhTri *tris;
    const ViewNode *vnodes;     // primary-ray mode: this frame's camera-space nodes (or null: walk the 3-D nodes)
    // primary-ray mode
    float cam[12]; // origin, U, V, W
    int width, height, x0, y0, w, h;
    // ray-buffer mode
    const float4 *rays;
    long long n_rays;
    // outputs
    float4 *hits;   // {t, id, u, v} or null
    uint32_t *bgra; // or null
    long long pitch_px;
    // shading inputs
    const float4 *pos, *nrm;
    const int *idx;
    cudaTextureObject_t tex;
    int tex_w, tex_h;
    int cull[4];   // primary mode: inclusive pixel rect [x0, y0, x1, y1] outside of which no ray can hit the scene
    int tt[4];     // primary mode: traced tile rectangle {tile x0, tile y0, tiles wide, tiles high} (8x4-pixel tiles of the rect)
    int trace_blocks; // primary mode: blocks [0, trace_blocks) trace 2x2 tiles each, the rest clear
    // image-space partition (primary mode): frame rows are grouped into stripes of 8 * st_R rows, a rank owns the stripes
    // s = rem (mod st_mod); st_mod = 1: everything.  Traced block rows (8 pixel rows each) are enumerated over owned stripes only:
    // the v-th owned block row of the frame is ((rem + (v / R) * mod) * R + v % R); st_v0 = index of the first one traced.
    unsigned st_R, st_mod, st_rem, st_v0;
    int st_fb_base; // frame block row of tile row tt[1]
    unsigned long long *stats; // optional: [0] inner-node visits, [1] triangle tests, [2] rays (instrumented build)
};

struct Hit { float t, u, v; unsigned id; };

template <bool STATS, bool FMA>
__device__ __forceinline__ Hit trace(const TraceArgs &a, float ox, float oy, float oz, float dx, float dy, float dz, int *stack /* [STACK], per thread */)
{
    const float ix = 1.0f / dx, iy = 1.0f / dy, iz = 1.0f / dz;
    // FMA form of the slab test, t = lo * inv - o * inv: half the FP32 instructions.  Only conservativeness matters for
    // box tests (hits are decided by the exact Moller-Trumbore below); the launcher enables it when the ray origin is
    // within 16 scene extents, where its rounding error (<= 2^-19 extent in box space) stays inside the 2^-17 padding.
    const float oix = ox * ix, oiy = oy * iy, oiz = oz * iz;
    unsigned long long best = ~0ull;
    float tbest = INFINITY, bu = 0.0f, bv = 0.0f;
    int sp = 0, cur = 0;
    unsigned n_nodes = 0, n_tests = 0;
    for (;;) {
        if (cur >= 0) {
            if (STATS) ++n_nodes;
            const float4 *np = reinterpret_cast<const float4 *>(a.nodes + cur);
            const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2);
            const int4 n3 = __ldg(reinterpret_cast<const int4 *>(np + 3));
            // fminf/fmaxf drop NaN (0 * inf on a slab face): conservative
            float a0, b0, c0, d0, e0, f0, a1, b1, c1, d1, e1, f1;
            if (FMA) {
                a0 = __fmaf_rn(n0.x, ix, -oix); b0 = __fmaf_rn(n0.y, ix, -oix); c0 = __fmaf_rn(n0.z, iy, -oiy); d0 = __fmaf_rn(n0.w, iy, -oiy);
                e0 = __fmaf_rn(n2.x, iz, -oiz); f0 = __fmaf_rn(n2.y, iz, -oiz);
                a1 = __fmaf_rn(n1.x, ix, -oix); b1 = __fmaf_rn(n1.y, ix, -oix); c1 = __fmaf_rn(n1.z, iy, -oiy); d1 = __fmaf_rn(n1.w, iy, -oiy);
                e1 = __fmaf_rn(n2.z, iz, -oiz); f1 = __fmaf_rn(n2.w, iz, -oiz);
            } else {
                a0 = (n0.x - ox) * ix; b0 = (n0.y - ox) * ix; c0 = (n0.z - oy) * iy; d0 = (n0.w - oy) * iy;
                e0 = (n2.x - oz) * iz; f0 = (n2.y - oz) * iz;
                a1 = (n1.x - ox) * ix; b1 = (n1.y - ox) * ix; c1 = (n1.z - oy) * iy; d1 = (n1.w - oy) * iy;
                e1 = (n2.z - oz) * iz; f1 = (n2.w - oz) * iz;
            }
            float tn0 = fmaxf(fmaxf(fminf(a0, b0), fminf(c0, d0)), fmaxf(fminf(e0, f0), 0.0f));
            float tf0 = fminf(fminf(fmaxf(a0, b0), fmaxf(c0, d0)), fminf(fmaxf(e0, f0), tbest));
            float tn1 = fmaxf(fmaxf(fminf(a1, b1), fminf(c1, d1)), fmaxf(fminf(e1, f1), 0.0f));
            float tf1 = fminf(fminf(fmaxf(a1, b1), fmaxf(c1, d1)), fminf(fmaxf(e1, f1), tbest));
            const bool h0 = tn0 <= tf0, h1 = tn1 <= tf1;
            if (h0 && h1) {
                const bool swap = tn1 < tn0;
                stack[sp] = swap ? n3.x : n3.y;
                ++sp;
                cur = swap ? n3.y : n3.x;
                continue;
            }
            if (h0) { cur = n3.x; continue; }
            if (h1) { cur = n3.y; continue; }
        } else {
            if (STATS) ++n_tests;
            const float4 *tp = reinterpret_cast<const float4 *>(a.tris + ~cur);
            const float4 v0 = __ldg(tp), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
            // Moller-Trumbore, operation for operation as oracle/raycast_oracle.c: rc_moller_trumbore
            const float px = dy * e2.z - dz * e2.y, py = dz * e2.x - dx * e2.z, pz = dx * e2.y - dy * e2.x;
            const float det = (e1.x * px + e1.y * py) + e1.z * pz;
            if (det != 0.0f) {
                const float inv = 1.0f / det;
                const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
                const float u = ((tx * px + ty * py) + tz * pz) * inv;
                if (u >= 0.0f && !(u > 1.0f)) {
                    const float qx = ty * e1.z - tz * e1.y, qy = tz * e1.x - tx * e1.z, qz = tx * e1.y - ty * e1.x;
                    const float v = ((dx * qx + dy * qy) + dz * qz) * inv;
                    if (v >= 0.0f && !(u + v > 1.0f)) {
                        const float t = ((e2.x * qx + e2.y * qy) + e2.z * qz) * inv;
                        if (t > 0.0f && t != INFINITY) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | __float_as_uint(v0.w);
                            if (key < best) { best = key; tbest = t; bu = u; bv = v; }
                        }
                    }
                }
            }
        }
        if (sp == 0) break;
        --sp;
        cur = stack[sp];
    }
    if (STATS) {
        atomicAdd(a.stats, (unsigned long long)n_nodes);
        atomicAdd(a.stats + 1, (unsigned long long)n_tests);
        atomicAdd(a.stats + 2, 1ull);
    }
    Hit h;
    h.t = tbest; h.u = bu; h.v = bv; h.id = best == ~0ull ? 0xFFFFFFFFu : (unsigned)best;
    return h;
}

// Traversal of this frame's ViewNodes: the 32 rays of a tile walk the hierarchy TOGETHER.  In screen space the traversal order
// (nearer c_min first) does not depend on the ray, so one warp-wide stack visits exactly the union of the nodes the lanes
// would visit on their own, each lane still culling with its own tbest through its vote.  Node and triangle fetches become
// one broadcast load per warp instead of up to 32 divergent ones, control flow is warp-uniform, and the stack (node,
// voters, c_min) lives in shared memory.  ALL lanes of the warp must call (lanes without a pixel pass live = false).
template <bool STATS>
__device__ __forceinline__ Hit trace_view_packet(const TraceArgs &a, bool live, float sx, float sy, float ox, float oy, float oz, float dx,
                                                 float dy, float dz, int4 *wstack /* [STACK], per warp, shared memory */)
{
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    unsigned long long best = ~0ull;
    float tbest = INFINITY, bu = 0.0f, bv = 0.0f;
    unsigned n_nodes = 0, n_tests = 0;
    int sp = 0, cur = 0;
    unsigned voters = __ballot_sync(FULL, live);
    while (voters != 0u) {
        if (cur >= 0) {
            if (STATS && ((voters >> lane) & 1u)) ++n_nodes;
            const float4 *np = reinterpret_cast<const float4 *>(a.vnodes + cur);
            const float4 r0 = __ldg(np), r1 = __ldg(np + 1), zc = __ldg(np + 2);
            const bool h0 = live && sx >= r0.x && sx <= r0.y && sy >= r0.z && sy <= r0.w && zc.x <= tbest;
            const bool h1 = live && sx >= r1.x && sx <= r1.y && sy >= r1.z && sy <= r1.w && zc.y <= tbest;
            const unsigned m0 = __ballot_sync(FULL, h0), m1 = __ballot_sync(FULL, h1);
            const int c0 = __float_as_int(zc.z), c1 = __float_as_int(zc.w);
            if (m0 != 0u && m1 != 0u) {
                const bool swap = zc.y < zc.x;
                if (lane == 0) wstack[sp] = make_int4(swap ? c0 : c1, (int)(swap ? m0 : m1), __float_as_int(swap ? zc.x : zc.y), 0);
                __syncwarp();
                ++sp;
                cur = swap ? c1 : c0; voters = swap ? m1 : m0;
                continue;
            }
            if (m0 != 0u) { cur = c0; voters = m0; continue; }
            if (m1 != 0u) { cur = c1; voters = m1; continue; }
        } else if ((voters >> lane) & 1u) {
            if (STATS) ++n_tests;
            const float4 *tp = reinterpret_cast<const float4 *>(a.tris + ~cur);
            const float4 v0 = __ldg(tp), e1 = __ldg(tp + 1), e2 = __ldg(tp + 2);
            // Moller-Trumbore, operation for operation as oracle/raycast_oracle.c: rc_moller_trumbore
            const float px = dy * e2.z - dz * e2.y, py = dz * e2.x - dx * e2.z, pz = dx * e2.y - dy * e2.x;
            const float det = (e1.x * px + e1.y * py) + e1.z * pz;
            if (det != 0.0f) {
                const float inv = 1.0f / det;
                const float tx = ox - v0.x, ty = oy - v0.y, tz = oz - v0.z;
                const float u = ((tx * px + ty * py) + tz * pz) * inv;
                if (u >= 0.0f && !(u > 1.0f)) {
                    const float qx = ty * e1.z - tz * e1.y, qy = tz * e1.x - tx * e1.z, qz = tx * e1.y - ty * e1.x;
                    const float v = ((dx * qx + dy * qy) + dz * qz) * inv;
                    if (v >= 0.0f && !(u + v > 1.0f)) {
                        const float t = ((e2.x * qx + e2.y * qy) + e2.z * qz) * inv;
                        if (t > 0.0f && t != INFINITY) {
                            const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | __float_as_uint(v0.w);
                            if (key < best) { best = key; tbest = t; bu = u; bv = v; }
                        }
                    }
                }
            }
        }
        // pop the nearest pending subtree some voter still needs (their tbest may have shrunk since the push)
        voters = 0u;
        while (sp > 0) {
            --sp;
            const int4 e = wstack[sp];
            const unsigned m = __ballot_sync(FULL, (((unsigned)e.y >> lane) & 1u) && __int_as_float(e.z) <= tbest);
            if (m != 0u) { cur = e.x; voters = m; break; }
        }
    }
    if (STATS && live) {
        atomicAdd(a.stats, (unsigned long long)n_nodes);
        atomicAdd(a.stats + 1, (unsigned long long)n_tests);
        atomicAdd(a.stats + 2, 1ull);
    }
    Hit h;
    h.t = tbest; h.u = bu; h.v = bv; h.id = best == ~0ull ? 0xFFFFFFFFu : (unsigned)best;
    return h;
}

// Lambert / texture shade of a hit, as oracle/raycast_oracle.c: orc_shade_hits
template <int SHADER>
__device__ __forceinline__ uint32_t shade(const TraceArgs &a, const Hit &h)
{
    if (h.id == 0xFFFFFFFFu) return 0u; // clear colour (0,0,0,0)
    long long i0 = 3ll * h.id, i1 = i0 + 1, i2 = i0 + 2;
    if (a.idx) { i0 = a.idx[i0]; i1 = a.idx[i1]; i2 = a.idx[i2]; }
    const float n = RT_INV_SQRT3;
    const float4 N0 = __ldg(a.nrm + i0), N1 = __ldg(a.nrm + i1), N2 = __ldg(a.nrm + i2);
    const float t0 = (N0.x * n + N0.y * n) + N0.z * n, t1 = (N1.x * n + N1.y * n) + N1.z * n, t2 = (N2.x * n + N2.y * n) + N2.z * n;
    const float w0 = 1.0f - h.u - h.v;
    if (SHADER == RT_SHADER_LESSON08) {
        const float d = fmaxf(0.2f, t0) * w0 + fmaxf(0.2f, t1) * h.u + fmaxf(0.2f, t2) * h.v;
        return rt_pack_bgra(d, d, d, 1.0f);
    }
    const float d = (0.2f + fmaxf(0.0f, t0)) * w0 + (0.2f + fmaxf(0.0f, t1)) * h.u + (0.2f + fmaxf(0.0f, t2)) * h.v;
    const float4 P0 = __ldg(a.pos + i0), P1 = __ldg(a.pos + i1), P2 = __ldg(a.pos + i2);
    const float cx = (P0.x * 2.0f) * w0 + (P1.x * 2.0f) * h.u + (P2.x * 2.0f) * h.v;
    const float cy = (P0.y * 2.0f) * w0 + (P1.y * 2.0f) * h.u + (P2.y * 2.0f) * h.v;
    const float4 tx = rt_sample2d(a.tex, a.tex_w, a.tex_h, cx, cy);
    return rt_pack_bgra(tx.x * d, tx.y * d, tx.z * d, 1.0f);
}

// Clear block: four pixel rows of the rect, minus the traced tiles, written as misses.
__device__ __forceinline__ void clear_band(const TraceArgs &a, int band)
{
    if (a.st_mod > 1u && ((unsigned)(a.y0 + band * 4) / (8u * a.st_R)) % a.st_mod != a.st_rem) return; // not this rank's stripe
    const bool traced_band = band >= a.tt[1] && band < a.tt[1] + a.tt[3];
    const int skip0 = traced_band ? a.tt[0] * 8 : a.w, skip1 = traced_band ? (a.tt[0] + a.tt[2]) * 8 : a.w;
    for (int r = 0; r < 4; ++r) {
        const int ly = band * 4 + r;
        if (ly >= a.h) break;
        for (int lx = threadIdx.x; lx < a.w; lx += TB) {
            if (lx >= skip0 && lx < skip1) continue;
            if (a.hits) a.hits[(long long)ly * a.w + lx] = make_float4(INFINITY, __uint_as_float(0xFFFFFFFFu), 0.0f, 0.0f);
            if (a.bgra) a.bgra[(long long)ly * a.pitch_px + lx] = 0u;
        }
    }
}

// MODE 0: rays from a buffer, hits out.  MODE 8 / 9: primary rays + shade with that lesson's shader.
//
// One warp per work unit (an 8x4-pixel tile or 32 consecutive rays), four units per block, and the hardware block
// scheduler as the load balancer.  The first version of this kernel was the textbook persistent-threads loop (resident
// warps claiming tiles with atomicAdd on one counter): at 4K that is 259 200 claims per frame, and the frame took exactly
// as long as the L2 needs to serialise them (1.25 ns per same-address atomic, 322 us) whatever the traversal cost -- the
// counter, not the BVH walk, was the bound.  Without it the same frame takes 279 us and per-visit savings show up again.
// Primary mode traces only the tiles that touch the caller's cull rectangle (blocks [0, trace_blocks), 2x2 tiles each);
// the rest of the pixel rect is written as misses by the clear blocks that follow them in the grid.
//
// Measured on B200 and NOT adopted (4K frame, dragon100k; all under the counter-bound loop, so only the losses are
// conclusive): claiming the next tile early (+6 %), claiming 8 tiles per atomic (+65 %: the tail grows), entry distances on
// the stack for pop-time culling (+10 %), per-lane refill a la Aila-Laine, idle lanes claiming single pixels (+27 %: primary
// rays are coherent, mixing tiles in a warp costs more in divergent node fetches than parked lanes do), a warp-synchronous
// while-while loop that parks leaves and runs node steps and triangle tests in separate ballot-driven phases (+44 %: lanes
// blocked on two parked leaves wait for the deepest lane of every phase).
// Measured after the counter was gone and NOT adopted (200 us with tile packets): per-lane stacks over the screen-space
// nodes (241 us: divergent node fetches, 17 of 32 lanes active); 16x8 .. 32x16-pixel region packets, a cell of 2x2 .. 4x4
// pixels per lane, hits in shared memory and the lanes re-dealt over each leaf's pixels (4x fewer node visits per ray, but
// 290 - 485 us: with ~6-pixel triangles the per-leaf work dominates and gets more expensive); prefetching both children of
// a node while the warp votes (+15 %: the loop is issue-bound, not latency-bound); 64- or 32-thread blocks (+5 - 10 %);
// the packet walk over the 3-D nodes with per-lane slab tests (100k triangles: 228 vs 253 us per-lane, but 1M: 512 vs 407 us,
// sub-pixel triangles leave nothing for a tile to share -- and below ~256k triangles the screen-space packets win anyway);
// a multiplicative permutation of the launch order, to keep the object's expensive tiles out of the kernel's tail (+-0 %);
// 32x4-pixel strips per block with colours swapped through shared memory so that every warp stores whole 128-byte rows,
// meant for frames in another GPU's memory (-5 % locally, +-0 % at N=4 over NVLink: store width is not what binds rank 0);
// a two-level walk (round 2, profiles/r02a_ab_experimental.txt): a block owns a region of 8x4 tiles, walks the top of the tree
// once for the region down to a frontier of nodes <= 2 / 8 / 32 tiles in area (shared memory), then its warps walk each tile
// from the frontier, nearest candidate first.  Hits identical, 3.6 instead of 18.7 node visits per tile below the frontier --
// and 299 / 298 / 319 us per cfg4 frame against 170 us (lesson08 camera at 4K: 389 - 413 against 408): the frontier scan
// (one entry per lane + a warp-min per candidate) and the block-wide breadth-first phase cost more than the repeated top
// visits, which are the cheap, fully coherent, L1-resident part of the walk.  The code was deleted.
template <int MODE, bool STATS, bool FMA, bool VIEW>
__global__ void __launch_bounds__(TB, RT_RAYCAST_MINB) raycast_kernel(const TraceArgs a)
{
    int stack[STACK];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (MODE) {
        if ((int)blockIdx.x >= a.trace_blocks) { clear_band(a, (int)blockIdx.x - a.trace_blocks); return; }
        const int bw = (a.tt[2] + TILES_BX - 1) / TILES_BX;
        int brow = (int)blockIdx.x / bw;
        if (a.st_mod > 1u) { // the brow-th traced block row of this rank -> its place in the frame
            const unsigned v = a.st_v0 + (unsigned)brow;
            brow = (int)((a.st_rem + (v / a.st_R) * a.st_mod) * a.st_R + v % a.st_R) - a.st_fb_base;
        }
        const int tx = a.tt[0] + TILES_BX * ((int)blockIdx.x % bw) + wid % TILES_BX, ty = a.tt[1] + TILES_BY * brow + wid / TILES_BX;
        if (tx >= a.tt[0] + a.tt[2] || ty >= a.tt[1] + a.tt[3]) return;
        const int lx = tx * 8 + (lane & 7), ly = ty * 4 + (lane >> 3);
        const bool live = lx < a.w && ly < a.h;
        const float sx = ((float)(a.x0 + lx) + 0.5f) * (2.0f / (float)a.width) - 1.0f;
        const float sy = 1.0f - ((float)(a.y0 + ly) + 0.5f) * (2.0f / (float)a.height);
        const float dx = (a.cam[3] * sx + a.cam[6] * sy) + a.cam[9];
        const float dy = (a.cam[4] * sx + a.cam[7] * sy) + a.cam[10];
        const float dz = (a.cam[5] * sx + a.cam[8] * sy) + a.cam[11];
        Hit h;
        if (VIEW) {
            __shared__ int4 wstacks[TB / 32][STACK];
            h = trace_view_packet<STATS>(a, live, sx, sy, a.cam[0], a.cam[1], a.cam[2], dx, dy, dz, wstacks[wid]);
            if (!live) return;
        } else {
            if (!live) return;
            h = trace<STATS, FMA>(a, a.cam[0], a.cam[1], a.cam[2], dx, dy, dz, stack);
        }
        if (a.hits) a.hits[(long long)ly * a.w + lx] = make_float4(h.t, __uint_as_float(h.id), h.u, h.v);
        if (a.bgra) a.bgra[(long long)ly * a.pitch_px + lx] = shade<MODE == 9 ? RT_SHADER_LESSON09 : RT_SHADER_LESSON08>(a, h);
    } else {
        const long long r = ((long long)blockIdx.x * (TB / 32) + wid) * 32 + lane;
        if (r >= a.n_rays) return;
        const float4 o = __ldg(a.rays + 2 * r), d = __ldg(a.rays + 2 * r + 1);
        const Hit h = trace<STATS, false>(a, o.x, o.y, o.z, d.x, d.y, d.z, stack);
        a.hits[r] = make_float4(h.t, __uint_as_float(h.id), h.u, h.v);
    }
}

template <int MODE, bool STATS, bool FMA, bool VIEW>
int launch_trace_s(TraceArgs &a, cudaStream_t st)
{
    long long blocks;
    if (MODE) {
        // traced tile rectangle: the 8x4 tiles of the pixel rect that touch the cull rectangle (frame pixels, inclusive)
        const int cx0 = (a.cull[0] > a.x0 ? a.cull[0] : a.x0) - a.x0, cy0 = (a.cull[1] > a.y0 ? a.cull[1] : a.y0) - a.y0;
        const int cx1 = (a.cull[2] < a.x0 + a.w - 1 ? a.cull[2] : a.x0 + a.w - 1) - a.x0;
        const int cy1 = (a.cull[3] < a.y0 + a.h - 1 ? a.cull[3] : a.y0 + a.h - 1) - a.y0;
        if (cx1 < cx0 || cy1 < cy0) { a.tt[0] = a.tt[1] = a.tt[2] = a.tt[3] = 0; }
        else { a.tt[0] = cx0 >> 3; a.tt[1] = cy0 >> 2; a.tt[2] = (cx1 >> 3) - a.tt[0] + 1; a.tt[3] = (cy1 >> 2) - a.tt[1] + 1; }
        int block_rows = (a.tt[3] + TILES_BY - 1) / TILES_BY;
        if (a.st_mod > 1u && a.tt[3] > 0) {
            // block rows must coincide with the frame's 8-row groups (y0 is a multiple of 8: checked by the caller)
            static_assert(TILES_BY == 2, "stripe partition: a block row is 8 pixel rows");
            const int t0 = a.tt[1] & ~1, t1 = a.tt[1] + a.tt[3]; // first traced tile row, rounded down to a block row
            a.tt[3] = t1 - t0; a.tt[1] = t0;
            a.st_fb_base = a.y0 / 8 + a.tt[1] / 2;
            const unsigned R = a.st_R, mod = a.st_mod, rem = a.st_rem;
            auto owned_below = [&](unsigned fb) { // owned block rows of the frame with index < fb
                const unsigned q = fb / R, r = fb % R;
                const unsigned full = q > rem ? (q - rem - 1u) / mod + 1u : 0u;
                return full * R + (q % mod == rem ? r : 0u);
            };
            const unsigned fb0 = (unsigned)a.st_fb_base, fb1 = fb0 + (unsigned)((a.tt[3] + 1) / 2);
            a.st_v0 = owned_below(fb0);
            block_rows = (int)(owned_below(fb1) - a.st_v0);
        }
        a.trace_blocks = ((a.tt[2] + TILES_BX - 1) / TILES_BX) * block_rows;
        const bool all_traced = a.st_mod == 1u && a.tt[0] == 0 && a.tt[1] == 0 && a.tt[2] * 8 >= a.w && a.tt[3] * 4 >= a.h;
        blocks = (long long)a.trace_blocks + (all_traced ? 0 : (a.h + 3) >> 2);
    } else {
        blocks = (a.n_rays + TB - 1) / TB;
    }
    if (blocks > 0) raycast_kernel<MODE, STATS, FMA, VIEW><<<(unsigned)blocks, TB, 0, st>>>(a);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

template <int MODE>
int launch_trace(TraceArgs &a, bool fma, cudaStream_t st)
{
    if constexpr (MODE != 0) {
        if (a.vnodes) return a.stats ? launch_trace_s<MODE, true, false, true>(a, st) : launch_trace_s<MODE, false, false, true>(a, st);
    }
    if (a.stats) return fma ? launch_trace_s<MODE, true, true, false>(a, st) : launch_trace_s<MODE, true, false, false>(a, st);
    return fma ? launch_trace_s<MODE, false, true, false>(a, st) : launch_trace_s<MODE, false, false, false>(a, st);
}

// [U V W]^-1 in double and the rectangle padding (see project_kernel); false if the camera basis is singular
bool camera_inverse(const float *cam, ProjectArgs &p)
{
    const double U[3] = {cam[3], cam[4], cam[5]}, V[3] = {cam[6], cam[7], cam[8]}, W[3] = {cam[9], cam[10], cam[11]};
    // M = [U V W] (columns); rows of the inverse are cross products over the determinant
    const double c0[3] = {V[1] * W[2] - V[2] * W[1], V[2] * W[0] - V[0] * W[2], V[0] * W[1] - V[1] * W[0]};
    const double c1[3] = {W[1] * U[2] - W[2] * U[1], W[2] * U[0] - W[0] * U[2], W[0] * U[1] - W[1] * U[0]};
    const double c2[3] = {U[1] * V[2] - U[2] * V[1], U[2] * V[0] - U[0] * V[2], U[0] * V[1] - U[1] * V[0]};
    const double det = U[0] * c0[0] + U[1] * c0[1] + U[2] * c0[2];
    if (!(det != 0.0) || !(det - det == 0.0)) return false;
    double rowsum = 0.0;
    for (int k = 0; k < 3; ++k) {
        p.minv[k] = c0[k] / det; p.minv[3 + k] = c1[k] / det; p.minv[6 + k] = c2[k] / det;
    }
    for (int r = 0; r < 3; ++r) {
        const double rs = fabs(p.minv[3 * r]) + fabs(p.minv[3 * r + 1]) + fabs(p.minv[3 * r + 2]);
        rowsum = rs > rowsum ? rs : rowsum;
        if (!(rs - rs == 0.0)) return false;
    }
    double dmax = 0.0; // bound of |direction component| over the frame, |sx|, |sy| <= 1
    for (int k = 0; k < 3; ++k) {
        const double d = fabs(U[k]) + fabs(V[k]) + fabs(W[k]);
        dmax = d > dmax ? d : dmax;
    }
    const double pad = 8.0 * rowsum * dmax * 2.384185791015625e-7; // 8 x rowsum x 2^-22 dmax
    if (!(pad < 1e-3)) return false; // ill-conditioned basis: not worth it, walk the 3-D nodes
    p.pad_s = (float)pad;
    p.rowsum = (float)(rowsum * 1.000001);
    p.o[0] = cam[0]; p.o[1] = cam[1]; p.o[2] = cam[2];
    return true;
}

} // namespace

extern "C" {

int rt_raycast_rays(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_rays, int64_t n_rays, void *d_hits,
                    void *stream)
{
    RT_REQUIRE(d_nodes && d_tris && n_triangles >= 1, "BVH");
    RT_REQUIRE(n_rays >= 0 && n_rays < (1ll << 36), "ray count");
    if (n_rays == 0) return RT_OK;
    RT_REQUIRE(d_rays && d_hits, "ray / hit buffers");
    RT_REQUIRE((((uintptr_t)d_rays | (uintptr_t)d_hits) & 15) == 0, "16-byte alignment");
    TraceArgs a = {};
    a.nodes = (const RtBvhNode *)d_nodes; a.tris = (const RtBvhTri *)d_tris;
    a.rays = (const float4 *)d_rays; a.n_rays = n_rays; a.hits = (float4 *)d_hits;
    return launch_trace<0>(a, false, (cudaStream_t)stream);
}

int rt_raycast_primary(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_pos4, const void *d_nrm4,
                       const int32_t *d_indices, const float *camera, int width, int height, int x0, int y0, int w, int h, int shader,
                       uint64_t tex_handle, void *d_hits, void *d_bgra, int64_t bgra_pitch_px, void *d_stats,
                       const int *cull_rect, int fast_slab, void *d_view_nodes, const int32_t *stripes, void *stream)
{
    RT_REQUIRE(d_nodes && d_tris && n_triangles >= 1, "BVH");
    RT_REQUIRE(camera, "camera");
    RT_REQUIRE(width > 0 && height > 0 && w >= 0 && h >= 0 && x0 >= 0 && y0 >= 0 && x0 + w <= width && y0 + h <= height, "pixel rect");
    RT_REQUIRE(shader == RT_SHADER_LESSON08 || shader == RT_SHADER_LESSON09, "shader id");
    RT_REQUIRE(!d_bgra || (d_nrm4 && bgra_pitch_px >= w), "shading needs normals and a pitch >= w");
    RT_REQUIRE(d_hits || d_bgra, "at least one output");
    if (w == 0 || h == 0) return RT_OK;
    TraceArgs a = {};
    a.nodes = (const RtBvhNode *)d_nodes; a.tris = (const RtBvhTri *)d_tris;
    for (int i = 0; i < 12; ++i) a.cam[i] = camera[i];
    a.width = width; a.height = height; a.x0 = x0; a.y0 = y0; a.w = w; a.h = h;
    a.hits = (float4 *)d_hits; a.bgra = (uint32_t *)d_bgra; a.pitch_px = bgra_pitch_px;
    a.pos = (const float4 *)d_pos4; a.nrm = (const float4 *)d_nrm4; a.idx = d_indices;
    a.stats = (unsigned long long *)d_stats;
    a.cull[0] = cull_rect ? cull_rect[0] : 0; a.cull[1] = cull_rect ? cull_rect[1] : 0;
    a.cull[2] = cull_rect ? cull_rect[2] : width - 1; a.cull[3] = cull_rect ? cull_rect[3] : height - 1;
    a.st_R = 1u; a.st_mod = 1u; a.st_rem = 0u; a.st_v0 = 0u; a.st_fb_base = 0;
    if (stripes && stripes[1] > 1) {
        RT_REQUIRE(stripes[0] >= 8 && stripes[0] % 8 == 0 && stripes[2] >= 0 && stripes[2] < stripes[1] && y0 % 8 == 0,
                   "stripes: {rows (multiple of 8), mod, 0 <= rem < mod}, and the pixel rect must start on a multiple of 8 rows");
        a.st_R = (unsigned)stripes[0] / 8u; a.st_mod = (unsigned)stripes[1]; a.st_rem = (unsigned)stripes[2];
    }
    if (d_view_nodes) {
        RT_REQUIRE(((uintptr_t)d_view_nodes & 15) == 0, "view nodes must be 16-byte aligned");
        ProjectArgs p = {};
        if (camera_inverse(camera, p)) {
            p.nodes = a.nodes; p.vnodes = (ViewNode *)d_view_nodes; p.n_inner = n_triangles > 1 ? n_triangles - 1 : 1;
            p.tris = a.tris; // leaf rectangles from the triangles themselves (189 -> 169 us per cfg4 frame against box rectangles)
            project_kernel<<<(unsigned)((p.n_inner + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p);
            RT_CUDA(cudaGetLastError());
            if (g_view_refit_passes > 0) {
                view_refit_kernel<<<(unsigned)((p.n_inner + 127) / 128), 128, 0, (cudaStream_t)stream>>>(p.vnodes, p.n_inner, g_view_refit_passes);
                RT_CUDA(cudaGetLastError());
            }
            a.vnodes = p.vnodes;
        }
    }
    if (shader == RT_SHADER_LESSON09) {
        RT_REQUIRE(!d_bgra || (tex_handle != 0 && d_pos4), "lesson09 shading needs a texture handle and positions");
        if (tex_handle) {
            const rt_texture *t = (const rt_texture *)(uintptr_t)tex_handle;
            a.tex = t->obj; a.tex_w = t->w; a.tex_h = t->h;
        }
        return launch_trace<9>(a, fast_slab != 0, (cudaStream_t)stream);
    }
    return launch_trace<8>(a, fast_slab != 0, (cudaStream_t)stream);
}

// Host only (no device work): {origin, U, V, W} (12 floats, model space) of the reference's camera convention
// (rendering/_core.py:528-548, row vectors: p_clip = ((p World) View) Proj, left-handed, w = z_view): the pixel centre with
// NDC coordinates (sx, sy) looks along U sx + V sy + W.  View = [R | 0; -eye R | 1] with the camera axes as COLUMNS of R,
// Proj scales x, y by proj[0][0], proj[1][1].  With a World matrix the frame is carried into model space with World^-1
// (general 4x4 inverse by cofactors, so scaled / sheared worlds work too).  All arithmetic in double, one rounding to float.
// Returns 1, or 0 when R or World is singular or the data is not finite.  A tutorial frame calls this once per frame; numpy
// needs ~140 us for the same arithmetic (two LAPACK inversions of tiny matrices), which is more than the 4K frame takes.
int rt_camera_frame(const float *view16, const float *proj16, const float *world16, float *out12)
{
    if (!view16 || !proj16 || !out12) return 0;
    double r[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r[i][j] = view16[4 * i + j];
    // inverse of R by cofactors
    const double c00 = r[1][1] * r[2][2] - r[1][2] * r[2][1], c01 = r[1][2] * r[2][0] - r[1][0] * r[2][2], c02 = r[1][0] * r[2][1] - r[1][1] * r[2][0];
    const double det = r[0][0] * c00 + r[0][1] * c01 + r[0][2] * c02;
    if (!(det != 0.0) || !(det - det == 0.0)) return 0;
    const double ri[3][3] = {{c00 / det, (r[0][2] * r[2][1] - r[0][1] * r[2][2]) / det, (r[0][1] * r[1][2] - r[0][2] * r[1][1]) / det},
                             {c01 / det, (r[0][0] * r[2][2] - r[0][2] * r[2][0]) / det, (r[0][2] * r[1][0] - r[0][0] * r[1][2]) / det},
                             {c02 / det, (r[0][1] * r[2][0] - r[0][0] * r[2][1]) / det, (r[0][0] * r[1][1] - r[0][1] * r[1][0]) / det}};
    const double t[3] = {view16[12], view16[13], view16[14]};
    double v[4][4]; // rows: eye (w = 1), U, V, W (w = 0)
    for (int j = 0; j < 3; ++j) v[0][j] = -(t[0] * ri[0][j] + t[1] * ri[1][j] + t[2] * ri[2][j]);
    v[0][3] = 1.0;
    const double ws = proj16[0], hs = proj16[5];
    for (int j = 0; j < 3; ++j) { v[1][j] = r[j][0] / ws; v[2][j] = r[j][1] / hs; v[3][j] = r[j][2]; }
    v[1][3] = v[2][3] = v[3][3] = 0.0;
    if (world16) {
        double m[16], inv[16];
        for (int i = 0; i < 16; ++i) m[i] = world16[i];
        inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
        inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
        inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
        inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
        inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
        inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
        inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
        inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
        inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
        inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
        inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
        inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
        inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
        inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
        inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
        inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
        const double d4 = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
        if (!(d4 != 0.0) || !(d4 - d4 == 0.0)) return 0;
        for (int k = 0; k < 4; ++k) {
            double o[3];
            for (int j = 0; j < 3; ++j) o[j] = (v[k][0] * inv[j] + v[k][1] * inv[4 + j] + v[k][2] * inv[8 + j] + v[k][3] * inv[12 + j]) / d4;
            v[k][0] = o[0]; v[k][1] = o[1]; v[k][2] = o[2];
        }
    }
    for (int k = 0; k < 4; ++k)
        for (int j = 0; j < 3; ++j) {
            if (!(v[k][j] - v[k][j] == 0.0)) return 0;
            out12[3 * k + j] = (float)v[k][j];
        }
    return 1;
}

// Host only (no device work): the cull rectangle callers hand to rt_raycast_primary.  Conservative inclusive pixel rect
// of the scene box [lo, hi] seen from `camera` (projected corners +- 2 px, clamped to the frame); returns 1 and fills
// rect, or 0 when there is no usable bound (a corner at or behind the eye plane, singular basis, non-finite data).
int rt_raycast_screen_bounds(const float *camera, const double *lo, const double *hi, int width, int height, int *rect)
{
    if (!camera || !lo || !hi || !rect || width <= 0 || height <= 0) return 0;
    const double o[3] = {camera[0], camera[1], camera[2]};
    const double U[3] = {camera[3], camera[4], camera[5]}, V[3] = {camera[6], camera[7], camera[8]}, W[3] = {camera[9], camera[10], camera[11]};
    const double r0[3] = {V[1] * W[2] - V[2] * W[1], V[2] * W[0] - V[0] * W[2], V[0] * W[1] - V[1] * W[0]};
    const double r1[3] = {W[1] * U[2] - W[2] * U[1], W[2] * U[0] - W[0] * U[2], W[0] * U[1] - W[1] * U[0]};
    const double r2[3] = {U[1] * V[2] - U[2] * V[1], U[2] * V[0] - U[0] * V[2], U[0] * V[1] - U[1] * V[0]};
    const double det = U[0] * r0[0] + U[1] * r0[1] + U[2] * r0[2];
    if (!(det != 0.0) || !(det - det == 0.0)) return 0;
    double extent = 0.0;
    for (int k = 0; k < 3; ++k) extent = hi[k] - lo[k] > extent ? hi[k] - lo[k] : extent;
    const double c_floor = 1e-6 * (extent > 1e-30 ? extent : 1e-30);
    double smin = INFINITY, smax = -INFINITY, tmin = INFINITY, tmax = -INFINITY;
    for (int k = 0; k < 8; ++k) {
        const double q[3] = {((k & 4) ? hi[0] : lo[0]) - o[0], ((k & 2) ? hi[1] : lo[1]) - o[1], ((k & 1) ? hi[2] : lo[2]) - o[2]};
        const double c = (r2[0] * q[0] + r2[1] * q[1] + r2[2] * q[2]) / det;
        if (!(c > c_floor) || !(c - c == 0.0)) return 0;
        const double s = (r0[0] * q[0] + r0[1] * q[1] + r0[2] * q[2]) / det / c, t = (r1[0] * q[0] + r1[1] * q[1] + r1[2] * q[2]) / det / c;
        if (!(s - s == 0.0) || !(t - t == 0.0)) return 0;
        smin = s < smin ? s : smin; smax = s > smax ? s : smax; tmin = t < tmin ? t : tmin; tmax = t > tmax ? t : tmax;
    }
    const double big = 1073741824.0; // keeps the int conversion defined for far off-screen boxes
    auto clampd = [&](double v) { return v < -big ? -big : (v > big ? big : v); };
    const double px0 = clampd(floor((smin + 1.0) * (width * 0.5) - 0.5) - 2.0), px1 = clampd(ceil((smax + 1.0) * (width * 0.5) - 0.5) + 2.0);
    const double py0 = clampd(floor((1.0 - tmax) * (height * 0.5) - 0.5) - 2.0), py1 = clampd(ceil((1.0 - tmin) * (height * 0.5) - 0.5) + 2.0);
    rect[0] = px0 > 0.0 ? (int)px0 : 0; rect[1] = py0 > 0.0 ? (int)py0 : 0;
    rect[2] = px1 < width - 1 ? (int)px1 : width - 1; rect[3] = py1 < height - 1 ? (int)py1 : height - 1;
    return 1;
}

// The same for n boxes (lo / hi: n x 3 doubles, e.g. the bounds of 64 chunks of the Morton-sorted leaves): the union of their
// rectangles, a quarter smaller than the rectangle of the scene's own box for the dragon orbit.  Returns 1 (x1 < x0: nothing
// on screen), or 0 as soon as one box has no usable bound.
int rt_raycast_screen_bounds_n(const float *camera, const double *lo, const double *hi, int n_boxes, int width, int height, int *rect)
{
    if (!rect || n_boxes < 1) return 0;
    int u[4] = {width, height, -1, -1}, r[4];
    for (int i = 0; i < n_boxes; ++i) {
        if (!rt_raycast_screen_bounds(camera, lo + 3 * i, hi + 3 * i, width, height, r)) return 0;
        if (r[2] < r[0] || r[3] < r[1]) continue;
        u[0] = r[0] < u[0] ? r[0] : u[0]; u[1] = r[1] < u[1] ? r[1] : u[1];
        u[2] = r[2] > u[2] ? r[2] : u[2]; u[3] = r[3] > u[3] ? r[3] : u[3];
    }
    for (int k = 0; k < 4; ++k) rect[k] = u[k];
    return 1;
}

// Number of tightening iterations (inside one launch of view_refit_kernel) rt_raycast_primary runs after the projection, 0..64.
// Process-wide.
int rt_raycast_set_view_refit(int passes)
{
    RT_REQUIRE(passes >= 0 && passes <= 64, "0..64 passes");
    g_view_refit_passes = passes;
    return RT_OK;
}

int64_t rt_raycast_view_node_bytes(int64_t n_triangles)
{
    return (int64_t)sizeof(ViewNode) * (n_triangles > 1 ? n_triangles - 1 : 1);
}

} // extern "C"
