// rt_raycast.cu -- Raycaster.ray_cast for sm_100a (rendering/_raycaster.py:35-36 is `pass` in the reference;
// semantics are those of oracle/raycast_oracle.c: float32 Moller-Trumbore without FMA, closest hit = min over
// (bits(t) << 32 | triangle id)).
//
// One persistent kernel: warps pull 8x4-pixel tiles (or 32-ray groups) from an atomic counter, every lane walks
// the LBVH with its own short stack (local memory, L1-resident), inner nodes are four 128-bit
// read-only loads that test both children, leaves are three.  Primary mode generates the ray from the camera
// frame in-kernel and shades the hit (Lambert / texture) straight into the BGRA8 frame, so per ray only 4 B
// (+16 B if hits are requested) leave the SM.
//
// Bound: FP32 pipe + L1/L2 latency (the BVH of a 100k-triangle mesh is ~11 MB, L2-resident); HBM sees only the
// frame.  No contraction anywhere, so no tensor cores.
#include "rt_common.cuh"
#include "rt_bvh.cuh"

namespace {

constexpr int TB = 128;   // threads per block
constexpr int STACK = 64; // Karras tree depth <= 64 (32 key bits + index tiebreak), one pending sibling per level
// The per-lane stack lives in local memory (L1-resident, only the touched depth is ever cached).  Measured on B200
// against a [depth][thread] shared-memory stack: 332 vs 374 us per 4K frame -- the 32 KB of shared memory per CTA
// cost more occupancy than the L1 round trips do.

struct TraceArgs {
    const RtBvhNode *nodes;
    const RtBvhTri *tris;
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
    unsigned *ctl; // [0] next work unit, [1] finished blocks; both zero between launches
    int cull[4];   // primary mode: inclusive pixel rect [x0, y0, x1, y1] outside of which no ray can hit the scene
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

// MODE 0: rays from a buffer, hits out.  MODE 8 / 9: primary rays + shade with that lesson's shader.
template <int MODE, bool STATS, bool FMA>
__global__ void __launch_bounds__(TB) raycast_kernel(const TraceArgs a)
{
    int stack[STACK];
    const int lane = threadIdx.x & 31;
    const unsigned FULL = 0xffffffffu;
    const int tiles_x = MODE ? (a.w + 7) >> 3 : 1;
    const long long n_units = MODE ? (long long)tiles_x * ((a.h + 3) >> 2) : (a.n_rays + 31) >> 5;
    const float two_over_w = MODE ? 2.0f / (float)a.width : 0.0f, two_over_h = MODE ? 2.0f / (float)a.height : 0.0f;

    // Measured on B200 and NOT adopted (4K frame, dragon100k, 322 us with this loop): claiming the next tile early (+6 %),
    // claiming 8 tiles per atomic (+65 %: the tail grows), entry distances on the stack for pop-time culling (+10 %),
    // per-lane refill a la Aila-Laine, idle lanes claiming single pixels (+27 %: primary rays are coherent, mixing tiles in
    // a warp costs more in divergent node fetches than parked lanes do), a warp-synchronous while-while loop that parks leaves
    // and runs node steps and triangle tests in separate ballot-driven phases (+44 %: lanes blocked on two parked leaves wait
    // for the deepest lane of every phase), 12 instead of 10 resident blocks per SM via __launch_bounds__ (40 registers, +-0 %).
    // The stall samples ncu books on this atomic are lanes waiting at the reconvergence point for the longest ray of their tile.
    for (;;) {
        unsigned unit = 0;
        if (lane == 0) unit = atomicAdd(a.ctl, 1u);
        unit = __shfl_sync(FULL, unit, 0);
        if ((long long)unit >= n_units) break;
        if (MODE) {
            const int tx = (int)(unit % (unsigned)tiles_x), ty = (int)(unit / (unsigned)tiles_x);
            const int lx = tx * 8 + (lane & 7), ly = ty * 4 + (lane >> 3);
            if (lx >= a.w || ly >= a.h) continue;
            const float sx = ((float)(a.x0 + lx) + 0.5f) * two_over_w - 1.0f;
            const float sy = 1.0f - ((float)(a.y0 + ly) + 0.5f) * two_over_h;
            const float dx = (a.cam[3] * sx + a.cam[6] * sy) + a.cam[9];
            const float dy = (a.cam[4] * sx + a.cam[7] * sy) + a.cam[10];
            const float dz = (a.cam[5] * sx + a.cam[8] * sy) + a.cam[11];
            Hit h;
            const int gx = a.x0 + lx, gy = a.y0 + ly;
            if (gx < a.cull[0] || gy < a.cull[1] || gx > a.cull[2] || gy > a.cull[3]) {
                h.t = INFINITY; h.u = 0.0f; h.v = 0.0f; h.id = 0xFFFFFFFFu; // outside the scene's screen bounds: a miss
            } else {
                h = trace<STATS, FMA>(a, a.cam[0], a.cam[1], a.cam[2], dx, dy, dz, stack);
            }
            const long long p = (long long)ly * a.w + lx;
            if (a.hits) a.hits[p] = make_float4(h.t, __uint_as_float(h.id), h.u, h.v);
            if (a.bgra) a.bgra[(long long)ly * a.pitch_px + lx] = shade<MODE == 9 ? RT_SHADER_LESSON09 : RT_SHADER_LESSON08>(a, h);
        } else {
            const long long r = (long long)unit * 32 + lane;
            if (r >= a.n_rays) continue;
            const float4 o = __ldg(a.rays + 2 * r), d = __ldg(a.rays + 2 * r + 1);
            const Hit h = trace<STATS, false>(a, o.x, o.y, o.z, d.x, d.y, d.z, stack);
            a.hits[r] = make_float4(h.t, __uint_as_float(h.id), h.u, h.v);
        }
    }
    // last block out re-arms the counters for the next launch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(a.ctl + 1, 1u) == gridDim.x - 1) { a.ctl[0] = 0u; a.ctl[1] = 0u; }
    }
}

template <int MODE, bool STATS, bool FMA>
int launch_trace_s(const TraceArgs &a, cudaStream_t st)
{
    static int per_sm = 0; // per instantiation; the occupancy query costs ~10 us of host time
    if (per_sm == 0) RT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, raycast_kernel<MODE, STATS, FMA>, TB, 0));
    raycast_kernel<MODE, STATS, FMA><<<rt_sm_count() * (per_sm > 0 ? per_sm : 1), TB, 0, st>>>(a);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

template <int MODE>
int launch_trace(const TraceArgs &a, bool fma, cudaStream_t st)
{
    if (a.stats) return fma ? launch_trace_s<MODE, true, true>(a, st) : launch_trace_s<MODE, true, false>(a, st);
    return fma ? launch_trace_s<MODE, false, true>(a, st) : launch_trace_s<MODE, false, false>(a, st);
}

} // namespace

extern "C" {

int rt_raycast_rays(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_rays, int64_t n_rays, void *d_hits,
                    void *d_ctl, void *stream)
{
    RT_REQUIRE(d_nodes && d_tris && n_triangles >= 1, "BVH");
    RT_REQUIRE(n_rays >= 0 && n_rays < (1ll << 36), "ray count");
    if (n_rays == 0) return RT_OK;
    RT_REQUIRE(d_rays && d_hits && d_ctl, "ray / hit / control buffers");
    RT_REQUIRE((((uintptr_t)d_rays | (uintptr_t)d_hits) & 15) == 0, "16-byte alignment");
    TraceArgs a = {};
    a.nodes = (const RtBvhNode *)d_nodes; a.tris = (const RtBvhTri *)d_tris;
    a.rays = (const float4 *)d_rays; a.n_rays = n_rays; a.hits = (float4 *)d_hits; a.ctl = (unsigned *)d_ctl;
    return launch_trace<0>(a, false, (cudaStream_t)stream);
}

int rt_raycast_primary(const void *d_nodes, const void *d_tris, int64_t n_triangles, const void *d_pos4, const void *d_nrm4,
                       const int32_t *d_indices, const float *camera, int width, int height, int x0, int y0, int w, int h, int shader,
                       uint64_t tex_handle, void *d_hits, void *d_bgra, int64_t bgra_pitch_px, void *d_ctl, void *d_stats,
                       const int *cull_rect, int fast_slab, void *stream)
{
    RT_REQUIRE(d_nodes && d_tris && n_triangles >= 1, "BVH");
    RT_REQUIRE(camera && d_ctl, "camera / control block");
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
    a.pos = (const float4 *)d_pos4; a.nrm = (const float4 *)d_nrm4; a.idx = d_indices; a.ctl = (unsigned *)d_ctl;
    a.stats = (unsigned long long *)d_stats;
    a.cull[0] = cull_rect ? cull_rect[0] : 0; a.cull[1] = cull_rect ? cull_rect[1] : 0;
    a.cull[2] = cull_rect ? cull_rect[2] : width - 1; a.cull[3] = cull_rect ? cull_rect[3] : height - 1;
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

} // extern "C"
