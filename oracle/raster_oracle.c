/*
 * oracle/raster_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference rasterizer (lleonart1984/rendertoy), stage by
 * stage, with the fragment stream materialised exactly as the reference does.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this; the product (rendertoy_b200/) never does.
 *
 * Arithmetic convention (the reference leaves it to the OpenCL compiler): strict
 * IEEE-754 binary32, evaluated left to right, no FMA contraction, correctly
 * rounded division.  Build with -O2 -ffp-contract=off -fno-fast-math.
 *
 * Pinning: checked against the reference's own kernel strings executed through
 * oracle/clshim (see oracle/README.md) -> tests/golden/: every depth word and colour byte of six
 * reference-run scenes is reproduced (tests/test_golden.py).
 *
 * Reference citations are relative to /root/reference/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_SHADER_LESSON08 8 /* tutorials/lesson08_rasterization.py:36-62 */
#define ORC_SHADER_LESSON09 9 /* tutorials/lesson09_texture_mapping.py:67-95 */
#define ORC_NO_WINNER 0xFFFFFFFFu

typedef struct {
    int shader;
    int width, height;
    const float *vs_globals; /* Transforms{World,View,Proj}: 48 floats, row-major (lesson08:19-23) */
    const float *tex;        /* lesson09 Materials.DiffuseMap texels, float4 per texel, row 0 first */
    int tex_w, tex_h;
} orc_config;

typedef struct {
    int64_t triangles_in;
    int64_t primitives;        /* visible_primitives, _raster.py:424 */
    int64_t skipped_z0;        /* primitives dropped by "already rendered" test, _raster.py:236; never counted as drawn:
                                  the reference's `while` at :428 then never terminates */
    int64_t dropped_large;     /* pixel_count >= 64*64, _raster.py:294 */
    int64_t fragments;         /* out_fragments, _raster.py:433 */
    int64_t fragments_offscreen; /* (int)proj.xy outside the target: reference behaviour undefined */
    int64_t tie_pixels;        /* pixels where >1 primitive produced the winning depth bits */
    int64_t pixels_written;
    int64_t over_capacity;     /* primitives whose pixel_count >= 32*W*H (two negative bbox extents multiply to a large
                                  positive count, :241): :245 never admits them, so the reference loops forever */
} orc_stats;

/* ---- vertex-out layout ------------------------------------------------------
 * lesson08 Vertex_Out {float4 proj; float3 C;}            32 B -> 8 floats, 7 used
 * lesson09 Vertex_Out {float4 proj; float3 L; float2 C;}  48 B -> 12 floats: proj 0-3, L 4-6, (pad 7), C 8-9
 */
static int orc_stride(int shader) { return shader == ORC_SHADER_LESSON08 ? 8 : 12; }

/* _core.py:86-88  mul(float4 v, float4x4 m): r_j = dot(v, column j) */
static inline void orc_mul(const float v[4], const float *m, float r[4])
{
    for (int j = 0; j < 4; ++j)
        r[j] = ((v[0] * m[j] + v[1] * m[4 + j]) + v[2] * m[8 + j]) + v[3] * m[12 + j];
}

/* normalize((float3)(1,1,1)).x : correctly rounded 1/sqrt(3) */
static const float ORC_INV_SQRT3 = 0.57735026918962576f;

/* lesson08:41-54 / lesson09:72-86.  in: MeshVertex as 20 floats (P@0 N@4 C@8 T@12 B@16, _modeling.py:22-28) */
static void orc_vertex_shader(int shader, const float *v, const float *g, float *o)
{
    const float n = ORC_INV_SQRT3;
    float dt = (v[4] * n + v[5] * n) + v[6] * n;
    float H[4] = {v[0], v[1], v[2], 1.0f}, T[4];
    orc_mul(H, g, T);       /* World */
    orc_mul(T, g + 16, H);  /* View  */
    orc_mul(H, g + 32, T);  /* Proj  */
    memset(o, 0, sizeof(float) * (size_t)orc_stride(shader));
    o[0] = T[0]; o[1] = T[1]; o[2] = T[2]; o[3] = T[3];
    if (shader == ORC_SHADER_LESSON08) {
        float d = fmaxf(0.2f, dt);
        o[4] = d; o[5] = d; o[6] = d;
    } else {
        float d = 0.2f + fmaxf(0.0f, dt);
        o[4] = d; o[5] = d; o[6] = d;
        o[8] = v[0] * 2.0f; o[9] = v[1] * 2.0f; /* o.C = vertex.P.xy * 2 */
    }
}

/* _raster.py:63-73 VertexProcess */
void orc_vertex_process(int shader, const float *mesh, int64_t n, const float *globals, float *out)
{
    const int st = orc_stride(shader);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i)
        orc_vertex_shader(shader, mesh + 20 * i, globals, out + st * i);
}

/* _raster.py:25-40 interpolate2: every field  v0*(1-alpha) + v1*alpha */
static inline void orc_lerp(const float *a, const float *b, float alpha, float *o, int st)
{
    const float om = 1.0f - alpha;
    for (int i = 0; i < st; ++i) o[i] = a[i] * om + b[i] * alpha;
}

/* _raster.py:152-205 TriangleAssembly + near clip.  Emits primitive id 2*t+k for the k-th output
 * triangle of input triangle t (the reference's order is an atomic race; ids make ours defined). */
static int orc_assemble_one(const float *vb, const int32_t *ib, int64_t t, int st, float *out /* 2*3*st */)
{
    int64_t i0 = ib ? ib[3 * t + 0] : 3 * t + 0;
    int64_t i1 = ib ? ib[3 * t + 1] : 3 * t + 1;
    int64_t i2 = ib ? ib[3 * t + 2] : 3 * t + 2;
    const float *v0 = vb + st * i0, *v1 = vb + st * i1, *v2 = vb + st * i2;
    float z0 = v0[2], z1 = v1[2], z2 = v2[2];
    int clip = (z0 < 0 ? 1 : 0) | (z1 < 0 ? 2 : 0) | (z2 < 0 ? 4 : 0);
    if (clip == 7) return 0;
    float v01[12], v12[12], v20[12];
    orc_lerp(v0, v1, -z0 / (z1 - z0), v01, st);
    orc_lerp(v1, v2, -z1 / (z2 - z1), v12, st);
    orc_lerp(v2, v0, -z2 / (z0 - z2), v20, st);
    const float *a, *b, *c;
    switch (clip) {
    case 0: a = v0;  b = v1;  c = v2;  break;
    case 1: a = v01; b = v1;  c = v2;  break;
    case 2: a = v0;  b = v01; c = v12; break;
    case 3: a = v12; b = v2;  c = v20; break;
    case 4: a = v0;  b = v1;  c = v12; break;
    case 5: a = v01; b = v1;  c = v12; break;
    default: a = v0; b = v01; c = v20; break; /* 6 */
    }
    memcpy(out, a, sizeof(float) * st); memcpy(out + st, b, sizeof(float) * st); memcpy(out + 2 * st, c, sizeof(float) * st);
    switch (clip) {
    case 1: a = v01; b = v2;  c = v20; break;
    case 2: a = v0;  b = v12; c = v2;  break;
    case 4: a = v0;  b = v12; c = v20; break;
    default: return 1;
    }
    out += 3 * st;
    memcpy(out, a, sizeof(float) * st); memcpy(out + st, b, sizeof(float) * st); memcpy(out + 2 * st, c, sizeof(float) * st);
    return 2;
}

/* _raster.py:118-133 Dehomogenize: xyz /= w; y *= -1; xy += 1; xy *= viewport*0.5 */
static inline void orc_dehomogenize(float *p, float half_w, float half_h)
{
    float w = p[3];
    p[0] = p[0] / w; p[1] = p[1] / w; p[2] = p[2] / w;
    p[1] = p[1] * -1.0f;
    p[0] = p[0] + 1.0f; p[1] = p[1] + 1.0f;
    p[0] = p[0] * half_w; p[1] = p[1] * half_h;
}

/* (int) of a float as GPUs do it: round toward zero, saturate, NaN -> 0.  (C leaves the
 * out-of-range case undefined; the reference ran on GPUs.) */
static inline int32_t orc_f2i(float f)
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

static inline uint32_t orc_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

typedef struct {
    float *frag;      /* stride floats each */
    uint32_t *prim;   /* primitive id of each fragment */
    uint32_t *cell;   /* row*W+col of the loop cell that produced it */
    int64_t n, cap;
} orc_fragbuf;

static void orc_fragbuf_push(orc_fragbuf *fb, const float *f, int st, uint32_t prim, uint32_t cell)
{
    if (fb->n == fb->cap) {
        fb->cap = fb->cap ? fb->cap * 2 : 4096;
        fb->frag = (float *)realloc(fb->frag, sizeof(float) * (size_t)st * (size_t)fb->cap);
        fb->prim = (uint32_t *)realloc(fb->prim, sizeof(uint32_t) * (size_t)fb->cap);
        fb->cell = (uint32_t *)realloc(fb->cell, sizeof(uint32_t) * (size_t)fb->cap);
    }
    memcpy(fb->frag + (size_t)st * fb->n, f, sizeof(float) * st);
    fb->prim[fb->n] = prim;
    fb->cell[fb->n] = cell;
    fb->n++;
}

/* _raster.py:227-327 TriangleRaster for one primitive (three dehomogenized vertices) */
static void orc_raster_one(const float *P, int st, int W, int H, uint32_t prim_id, orc_fragbuf *fb,
                           int64_t *skipped_z0, int64_t *dropped_large, int64_t *over_capacity)
{
    const float *v1 = P, *v2 = P + st, *v3 = P + 2 * st;
    if (v1[2] < 0) { (*skipped_z0)++; return; } /* :236 "already rendered" */
    int64_t startx = orc_f2i(fminf(v1[0], fminf(v2[0], v3[0]))); if (startx < 0) startx = 0;
    int64_t starty = orc_f2i(fminf(v1[1], fminf(v2[1], v3[1]))); if (starty < 0) starty = 0;
    int64_t endx = 1 + (int64_t)orc_f2i(fmaxf(v1[0], fmaxf(v2[0], v3[0]))); if (endx > W - 1) endx = W - 1;
    int64_t endy = 1 + (int64_t)orc_f2i(fmaxf(v1[1], fmaxf(v2[1], v3[1]))); if (endy > H - 1) endy = H - 1;
    int64_t pixel_count = (endx - startx + 1) * (endy - starty + 1);
    if (pixel_count >= 32ll * W * H) (*over_capacity)++;

    float ax = v1[0], ay = v1[1], bx = v2[0], by = v2[1], cx = v3[0], cy = v3[1];
    float e1x = bx - ax, e1y = by - ay, e2x = cx - ax, e2y = cy - ay;
    int ccw = (e1x * e2y - e1y * e2x) <= 0;
    if (!ccw) { const float *t = v2; v2 = v3; v3 = t; }
    const float *h1 = v1, *h2 = v2, *h3 = v3;

    float a1 = h2[1] - h1[1], b1 = h1[0] - h2[0], c1 = h1[0] * (h1[1] - h2[1]) - h1[1] * (h1[0] - h2[0]);
    float a2 = h3[1] - h2[1], b2 = h2[0] - h3[0], c2 = h2[0] * (h2[1] - h3[1]) - h2[1] * (h2[0] - h3[0]);
    float a3 = h1[1] - h3[1], b3 = h3[0] - h1[0], c3 = h3[0] * (h3[1] - h1[1]) - h3[1] * (h3[0] - h1[0]);

    int t12 = (h1[1] == h2[1] && h2[0] <= h1[0]) || h1[1] < h2[1];
    int t23 = (h2[1] == h3[1] && h3[0] <= h2[0]) || h2[1] < h3[1];
    int t31 = (h3[1] == h1[1] && h1[0] <= h3[0]) || h3[1] < h1[1];
    const float eps = (float)0.00000001;
    float comp3 = t12 ? 0.0f : eps, comp1 = t23 ? 0.0f : eps, comp2 = t31 ? 0.0f : eps;

    if (!(pixel_count < 64 * 64)) { if (endx >= startx && endy >= starty) (*dropped_large)++; return; }
    float frag[12];
    for (int64_t row = starty; row <= endy; ++row)
        for (int64_t col = startx; col <= endx; ++col) {
            float px = (float)col + 0.5f, py = (float)row + 0.5f;
            float d1 = a1 * px + b1 * py + c1;
            float d2 = a2 * px + b2 * py + c2;
            float d3 = a3 * px + b3 * py + c3;
            float s = d1 + d2 + d3;
            float alpha3 = d1 / s, alpha1 = d2 / s, alpha2 = d3 / s;
            if (alpha1 >= comp1 && alpha2 >= comp2 && alpha3 >= comp3) {
                float q1 = alpha1 / h1[3], q2 = alpha2 / h2[3], q3 = alpha3 / h3[3];
                float qs = q1 + q2 + q3;
                float beta2 = q2 / qs, beta3 = q3 / qs;
                float w1 = 1.0f - beta2 - beta3;
                for (int i = 4; i < st; ++i) frag[i] = v1[i] * w1 + v2[i] * beta2 + v3[i] * beta3;
                for (int i = 0; i < 4; ++i) frag[i] = h1[i] * alpha1 + h2[i] * alpha2 + h3[i] * alpha3;
                orc_fragbuf_push(fb, frag, st, prim_id, (uint32_t)(row * W + col));
            }
        }
}

static inline void orc_atomic_min_u32(uint32_t *p, uint32_t v)
{
    uint32_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
}
static inline void orc_atomic_min_u64(uint64_t *p, uint64_t v)
{
    uint64_t cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, 1, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
}

/* _core.py:94-96 wrap_coord / sample2D (nearest texel, repeat).  Index clamped to the texture
 * (the reference would read the neighbouring pool bytes when rounding yields width/height). */
static inline float orc_wrap(float c) { return fmodf(fmodf(c, 1.0f) + 1.0f, 1.0f); }
static inline void orc_sample2d(const orc_config *cfg, float cx, float cy, float out[4])
{
    int32_t row = orc_f2i(orc_wrap(cy) * (float)cfg->tex_h);
    int32_t col = orc_f2i(orc_wrap(cx) * (float)cfg->tex_w);
    if (row < 0) row = 0; if (row > cfg->tex_h - 1) row = cfg->tex_h - 1;
    if (col < 0) col = 0; if (col > cfg->tex_w - 1) col = cfg->tex_w - 1;
    const float *t = cfg->tex + 4 * ((size_t)row * cfg->tex_w + col);
    out[0] = t[0]; out[1] = t[1]; out[2] = t[2]; out[3] = t[3];
}

/* lesson08:58-62 / lesson09:90-95 fragment shaders */
static inline void orc_fragment_shader(const orc_config *cfg, const float *f, float color[4])
{
    if (cfg->shader == ORC_SHADER_LESSON08) {
        color[0] = f[4]; color[1] = f[5]; color[2] = f[6]; color[3] = 1.0f;
    } else {
        float t[4];
        orc_sample2d(cfg, f[8], f[9], t);
        color[0] = t[0] * f[4]; color[1] = t[1] * f[5]; color[2] = t[2] * f[6]; color[3] = 1.0f;
    }
}

/* write_imagef to a CL_BGRA / CL_UNORM_INT8 image (_core.py:340): sat + round-to-nearest-even */
static inline uint8_t orc_unorm8(float c)
{
    float v = c * 255.0f;
    if (!(v > 0.0f)) return 0; /* also NaN */
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)nearbyintf(v);
}
void orc_pack_bgra(const float color[4], uint8_t out[4])
{
    out[0] = orc_unorm8(color[2]); out[1] = orc_unorm8(color[1]);
    out[2] = orc_unorm8(color[0]); out[3] = orc_unorm8(color[3]);
}

/* Raster.draw_triangles, _raster.py:416-437 (single-pass form: the multi-pass capacity loop only
 * changes batching).  depth/bgra accumulate across calls like the reference's persistent targets.
 * winner[p] = primitive id that owns pixel p after THIS draw, or ORC_NO_WINNER if untouched. */
int orc_draw_triangles(const orc_config *cfg, const float *mesh_vertices, const int32_t *indices, int64_t n_tris,
                       int64_t n_vertices, uint32_t *depth, uint8_t *bgra, uint32_t *winner, uint8_t *tie_mask,
                       orc_stats *stats)
{
    const int st = orc_stride(cfg->shader), W = cfg->width, H = cfg->height;
    orc_stats S; memset(&S, 0, sizeof S);
    S.triangles_in = n_tris;
    if (n_vertices <= 0) n_vertices = 3 * n_tris;

    /* 1. VertexProcess over every vertex of the buffer (:421 launches 3*T threads on the soup) */
    float *vb = (float *)malloc(sizeof(float) * (size_t)st * (size_t)(n_vertices > 0 ? n_vertices : 1));
    orc_vertex_process(cfg->shader, mesh_vertices, n_vertices, cfg->vs_globals, vb);

    /* 2. TriangleAssembly (:423) -> count, prefix, fill (deterministic stand-in for atomic_add order) */
    int32_t *cnt = (int32_t *)calloc((size_t)n_tris + 1, sizeof(int32_t));
    float *tmp_all = (float *)malloc(sizeof(float) * 6 * (size_t)st * (size_t)(n_tris > 0 ? n_tris : 1));
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_tris; ++t)
        cnt[t] = orc_assemble_one(vb, indices, t, st, tmp_all + 6 * (size_t)st * t);
    int64_t nprim = 0;
    int64_t *off = (int64_t *)malloc(sizeof(int64_t) * ((size_t)n_tris + 1));
    for (int64_t t = 0; t < n_tris; ++t) { off[t] = nprim; nprim += cnt[t]; }
    S.primitives = nprim;
    float *prims = (float *)malloc(sizeof(float) * 3 * (size_t)st * (size_t)(nprim > 0 ? nprim : 1));
    uint32_t *prim_id = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(nprim > 0 ? nprim : 1));
#pragma omp parallel for schedule(static)
    for (int64_t t = 0; t < n_tris; ++t)
        for (int k = 0; k < cnt[t]; ++k) {
            memcpy(prims + 3 * (size_t)st * (off[t] + k), tmp_all + 6 * (size_t)st * t + 3 * (size_t)st * k,
                   sizeof(float) * 3 * st);
            prim_id[off[t] + k] = (uint32_t)(2 * t + k);
        }
    free(tmp_all); free(cnt); free(off); free(vb);

    /* 3. Dehomogenize (:425) */
    const float half_w = (float)W * 0.5f, half_h = (float)H * 0.5f;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < 3 * nprim; ++i) orc_dehomogenize(prims + (size_t)st * i, half_w, half_h);

    /* 4. TriangleRaster (:431): per-thread fragment streams */
    int nth = 1;
#ifdef _OPENMP
    nth = omp_get_max_threads();
#endif
    orc_fragbuf *fbs = (orc_fragbuf *)calloc((size_t)nth, sizeof(orc_fragbuf));
    int64_t skipped = 0, dropped = 0, overcap = 0;
#pragma omp parallel reduction(+ : skipped, dropped, overcap)
    {
        int tid = 0;
#ifdef _OPENMP
        tid = omp_get_thread_num();
#endif
#pragma omp for schedule(dynamic, 256)
        for (int64_t p = 0; p < nprim; ++p)
            orc_raster_one(prims + 3 * (size_t)st * p, st, W, H, prim_id[p], &fbs[tid], &skipped, &dropped, &overcap);
    }
    S.skipped_z0 = skipped; S.dropped_large = dropped; S.over_capacity = overcap;
    for (int t = 0; t < nth; ++t) S.fragments += fbs[t].n;

    /* 5. DepthTest (:434, kernel :80-93) */
    int64_t offscreen = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : offscreen)
    for (int t = 0; t < nth; ++t)
        for (int64_t i = 0; i < fbs[t].n; ++i) {
            const float *f = fbs[t].frag + (size_t)st * i;
            if (f[2] < 0) continue;
            int32_t px = orc_f2i(f[0]), py = orc_f2i(f[1]);
            if (px < 0 || px >= W || py < 0 || py >= H) { offscreen++; continue; }
            orc_atomic_min_u32(depth + (size_t)py * W + px, orc_bits(f[2]));
        }
    S.fragments_offscreen = offscreen;

    /* 6. FragmentProcess (:436, kernel :95-112).  The reference lets every fragment whose depth bits
     * equal the final depth write (last writer wins, a race); we pick min (primitive id, cell rank). */
    uint64_t *sel = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)W * H);
    memset(sel, 0xFF, sizeof(uint64_t) * (size_t)W * H);
#pragma omp parallel for schedule(static, 1)
    for (int t = 0; t < nth; ++t)
        for (int64_t i = 0; i < fbs[t].n; ++i) {
            const float *f = fbs[t].frag + (size_t)st * i;
            if (f[2] < 0) continue;
            int32_t px = orc_f2i(f[0]), py = orc_f2i(f[1]);
            if (px < 0 || px >= W || py < 0 || py >= H) continue;
            size_t p = (size_t)py * W + px;
            if (depth[p] != orc_bits(f[2])) continue;
            uint32_t rank = fbs[t].cell[i] == (uint32_t)p ? 0u : 1u + fbs[t].cell[i];
            orc_atomic_min_u64(sel + p, ((uint64_t)fbs[t].prim[i] << 32) | rank);
        }
    int64_t written = 0;
#pragma omp parallel for schedule(static, 1) reduction(+ : written)
    for (int t = 0; t < nth; ++t)
        for (int64_t i = 0; i < fbs[t].n; ++i) {
            const float *f = fbs[t].frag + (size_t)st * i;
            float color[4];
            orc_fragment_shader(cfg, f, color); /* :100 runs before the tests */
            if (f[2] < 0) continue;             /* never entered the depth buffer */
            int32_t px = orc_f2i(f[0]), py = orc_f2i(f[1]);
            if (px < 0 || px >= W || py < 0 || py >= H) continue;
            size_t p = (size_t)py * W + px;
            if (depth[p] != orc_bits(f[2])) continue;
            uint32_t rank = fbs[t].cell[i] == (uint32_t)p ? 0u : 1u + fbs[t].cell[i];
            uint64_t me = ((uint64_t)fbs[t].prim[i] << 32) | rank;
            if (me != sel[p]) { if (tie_mask && (uint32_t)(sel[p] >> 32) != fbs[t].prim[i]) tie_mask[p] = 1; continue; }
            if (winner) winner[p] = fbs[t].prim[i];
            if (f[2] <= 0) continue;            /* :102 */
            orc_pack_bgra(color, bgra + 4 * p);
            written++;
        }
    S.pixels_written = written;
    if (tie_mask) { int64_t n = 0; for (size_t p = 0; p < (size_t)W * H; ++p) n += tie_mask[p]; S.tie_pixels = n; }

    for (int t = 0; t < nth; ++t) { free(fbs[t].frag); free(fbs[t].prim); free(fbs[t].cell); }
    free(fbs); free(sel); free(prims); free(prim_id);
    if (stats) *stats = S;
    return 0;
}

/* Raster.draw_points, _raster.py:399-414: VertexProcess -> PointAssembly (z<0 cull, :140-151) -> PointRaster
 * (|x|,|y| <= w clip on CLIP-space coordinates, :214-226) -> Dehomogenize -> DepthTest -> FragmentProcess.
 * Point id = position in the (optionally indexed) point list; ties go to the lowest id.  Fragments landing on
 * px == W or py == H (x == w exactly) are dropped (the reference would write out of bounds). */
int orc_draw_points(const orc_config *cfg, const float *mesh_vertices, const int32_t *indices, int64_t n_points, int64_t n_vertices,
                    uint32_t *depth, uint8_t *bgra, uint32_t *winner, orc_stats *stats)
{
    const int st = orc_stride(cfg->shader), W = cfg->width, H = cfg->height;
    orc_stats S; memset(&S, 0, sizeof S);
    S.triangles_in = n_points;
    float *vb = (float *)malloc(sizeof(float) * (size_t)st * (size_t)(n_vertices > 0 ? n_vertices : 1));
    orc_vertex_process(cfg->shader, mesh_vertices, n_vertices, cfg->vs_globals, vb);
    float *frag = (float *)malloc(sizeof(float) * (size_t)st * (size_t)(n_points > 0 ? n_points : 1));
    uint8_t *alive = (uint8_t *)calloc((size_t)(n_points > 0 ? n_points : 1), 1);
    const float half_w = (float)W * 0.5f, half_h = (float)H * 0.5f;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n_points; ++i) {
        const float *v = vb + (size_t)st * (indices ? indices[i] : i);
        if (v[2] < 0) continue;                                                   /* PointAssembly */
        if (v[0] < -v[3] || v[0] > v[3] || v[1] < -v[3] || v[1] > v[3]) continue; /* PointRaster   */
        float *f = frag + (size_t)st * i;
        memcpy(f, v, sizeof(float) * st);
        orc_dehomogenize(f, half_w, half_h);
        alive[i] = 1;
    }
    int64_t nfrag = 0, off = 0;
    for (int64_t i = 0; i < n_points; ++i) {
        if (!alive[i]) continue;
        nfrag++;
        const float *f = frag + (size_t)st * i;
        if (f[2] < 0) continue;
        int32_t px = orc_f2i(f[0]), py = orc_f2i(f[1]);
        if (px < 0 || px >= W || py < 0 || py >= H) { off++; continue; }
        uint32_t *d = depth + (size_t)py * W + px;
        if (orc_bits(f[2]) < *d) *d = orc_bits(f[2]);
    }
    uint32_t *sel = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)W * H);
    memset(sel, 0xFF, sizeof(uint32_t) * (size_t)W * H);
    for (int64_t i = 0; i < n_points; ++i) {
        if (!alive[i]) continue;
        const float *f = frag + (size_t)st * i;
        if (f[2] < 0) continue;
        int32_t px = orc_f2i(f[0]), py = orc_f2i(f[1]);
        if (px < 0 || px >= W || py < 0 || py >= H) continue;
        size_t p = (size_t)py * W + px;
        if (depth[p] == orc_bits(f[2]) && (uint32_t)i < sel[p]) sel[p] = (uint32_t)i;
    }
    int64_t written = 0;
    for (size_t p = 0; p < (size_t)W * H; ++p) {
        if (sel[p] == ORC_NO_WINNER) continue;
        const float *f = frag + (size_t)st * sel[p];
        float color[4];
        orc_fragment_shader(cfg, f, color);
        if (winner) winner[p] = sel[p];
        if (f[2] <= 0) continue;
        orc_pack_bgra(color, bgra + 4 * p);
        written++;
    }
    S.primitives = nfrag; S.fragments = nfrag; S.fragments_offscreen = off; S.pixels_written = written;
    free(vb); free(frag); free(alive); free(sel);
    if (stats) *stats = S;
    return 0;
}

/* Single-vertex known-answer helper: clip-space H and the dehomogenized proj */
void orc_vertex_kat(const float P[3], const float *globals, int W, int H, float clip[4], float screen[4])
{
    float v[20] = {0}, o[12];
    v[0] = P[0]; v[1] = P[1]; v[2] = P[2];
    orc_vertex_shader(ORC_SHADER_LESSON08, v, globals, o);
    memcpy(clip, o, 16);
    orc_dehomogenize(o, (float)W * 0.5f, (float)H * 0.5f);
    memcpy(screen, o, 16);
}

/* _core.py:376-388 clear(): fill depth with the bits of a float, colour with rgba (packed BGRA8) */
void orc_clear_depth(uint32_t *depth, int64_t n, float value)
{
    uint32_t b = orc_bits(value);
    for (int64_t i = 0; i < n; ++i) depth[i] = b;
}
void orc_clear_color(uint8_t *bgra, int64_t n, const float rgba[4])
{
    uint8_t px[4]; orc_pack_bgra(rgba, px);
    for (int64_t i = 0; i < n; ++i) memcpy(bgra + 4 * i, px, 4);
}

/* bench.py's CPU legs: torchrun exports OMP_NUM_THREADS=1 to its workers; the baseline is "all host cores", set explicitly */
void orc_set_threads(int n)
{
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
