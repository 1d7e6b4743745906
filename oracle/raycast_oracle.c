/*
 * oracle/raycast_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * PARITY UNPINNED: the reference has no ray caster (rendering/_raycaster.py:25-36 is a stub whose
 * ray_cast body is `pass`), so there is nothing to restate and no reference output to pin against.
 * This file DEFINES the closest-hit semantics the CUDA path is held to:
 *   - float32 Moller-Trumbore, evaluated left to right, no FMA, two-sided (the rasterizer does not cull),
 *   - closest hit = minimum of the 64-bit key (bits(t) << 32 | triangle id), t > 0,
 *   - primary rays from the reference camera convention (rendering/_core.py:528-548),
 *   - Lambert term of tutorials/lesson08_rasterization.py:42 / lesson09:73 interpolated with (u, v).
 * orc_raycast_brute is the definition; orc_bvh_* is a CPU BVH used for full-size checks and as the
 * timed CPU baseline, validated against the brute force in tests/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_MISS_ID 0xFFFFFFFFu

static inline uint32_t rc_bits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline int32_t rc_f2i(float f) /* GPU-style (int): truncate, saturate, NaN -> 0 */
{
    if (f != f) return 0;
    if (f >= 2147483648.0f) return INT32_MAX;
    if (f <= -2147483648.0f) return INT32_MIN;
    return (int32_t)f;
}

/* mesh: MeshVertex rows of 20 floats (rendering/_modeling.py:22-28); indices optional (int32) */
static inline const float *rc_vertex(const float *mesh, const int32_t *ib, int64_t t, int k)
{
    int64_t i = ib ? ib[3 * t + k] : 3 * t + k;
    return mesh + 20 * i;
}

/* returns 1 on hit, writing t,u,v */
static inline int rc_moller_trumbore(const float *o, const float *d, const float *v0, const float *v1,
                                     const float *v2, float *t_out, float *u_out, float *v_out)
{
    float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
    float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
    float px = d[1] * e2z - d[2] * e2y, py = d[2] * e2x - d[0] * e2z, pz = d[0] * e2y - d[1] * e2x;
    float det = (e1x * px + e1y * py) + e1z * pz;
    if (det == 0.0f) return 0;
    float inv = 1.0f / det;
    float tx = o[0] - v0[0], ty = o[1] - v0[1], tz = o[2] - v0[2];
    float u = ((tx * px + ty * py) + tz * pz) * inv;
    if (!(u >= 0.0f) || u > 1.0f) return 0;
    float qx = ty * e1z - tz * e1y, qy = tz * e1x - tx * e1z, qz = tx * e1y - ty * e1x;
    float v = ((d[0] * qx + d[1] * qy) + d[2] * qz) * inv;
    if (!(v >= 0.0f) || u + v > 1.0f) return 0;
    float t = ((e2x * qx + e2y * qy) + e2z * qz) * inv;
    if (!(t > 0.0f) || t == INFINITY) return 0;
    *t_out = t; *u_out = u; *v_out = v;
    return 1;
}

/* rays: N x {ox,oy,oz,_, dx,dy,dz,_} (two float3, 32 B, OpenCL float3 padding) */
void orc_raycast_brute(const float *mesh, const int32_t *ib, int64_t n_tris, const float *rays, int64_t n_rays,
                       float *out_t, uint32_t *out_id, float *out_u, float *out_v)
{
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float *o = rays + 8 * r, *d = rays + 8 * r + 4;
        uint64_t best = ~0ull; float bu = 0, bv = 0;
        for (int64_t t = 0; t < n_tris; ++t) {
            float tt, u, v;
            if (rc_moller_trumbore(o, d, rc_vertex(mesh, ib, t, 0), rc_vertex(mesh, ib, t, 1), rc_vertex(mesh, ib, t, 2), &tt, &u, &v)) {
                uint64_t key = ((uint64_t)rc_bits(tt) << 32) | (uint32_t)t;
                if (key < best) { best = key; bu = u; bv = v; }
            }
        }
        if (best == ~0ull) { out_t[r] = INFINITY; out_id[r] = ORC_MISS_ID; out_u[r] = 0; out_v[r] = 0; }
        else { uint32_t b = (uint32_t)(best >> 32); memcpy(out_t + r, &b, 4); out_id[r] = (uint32_t)best; out_u[r] = bu; out_v[r] = bv; }
    }
}

/* ---- CPU BVH (median split, <= 4 triangles per leaf), conservative boxes --------------------- */
typedef struct { float lo[3], hi[3]; int32_t left, right, first, count; } rc_node;
typedef struct {
    rc_node *nodes; int64_t n_nodes, cap_nodes;
    uint32_t *order; /* triangle ids in leaf order */
    float *tri;      /* 9 floats per triangle in ORIGINAL id order */
    float *cent;
    int64_t n_tris;
} rc_bvh;

static int rc_axis; static const float *rc_cent;
static int rc_cmp(const void *a, const void *b)
{
    float ca = rc_cent[3 * (size_t)(*(const uint32_t *)a) + rc_axis], cb = rc_cent[3 * (size_t)(*(const uint32_t *)b) + rc_axis];
    return (ca > cb) - (ca < cb);
}

static int32_t rc_build_rec(rc_bvh *b, int64_t first, int64_t count, float pad)
{
    int32_t me = (int32_t)b->n_nodes++;
    rc_node *n = &b->nodes[me];
    for (int a = 0; a < 3; ++a) { n->lo[a] = INFINITY; n->hi[a] = -INFINITY; }
    float clo[3] = {INFINITY, INFINITY, INFINITY}, chi[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = first; i < first + count; ++i) {
        const float *t = b->tri + 9 * (size_t)b->order[i];
        for (int k = 0; k < 3; ++k)
            for (int a = 0; a < 3; ++a) { n->lo[a] = fminf(n->lo[a], t[3 * k + a]); n->hi[a] = fmaxf(n->hi[a], t[3 * k + a]); }
        const float *c = b->cent + 3 * (size_t)b->order[i];
        for (int a = 0; a < 3; ++a) { clo[a] = fminf(clo[a], c[a]); chi[a] = fmaxf(chi[a], c[a]); }
    }
    for (int a = 0; a < 3; ++a) { n->lo[a] -= pad; n->hi[a] += pad; }
    n->first = (int32_t)first; n->count = (int32_t)count; n->left = n->right = -1;
    if (count <= 4) return me;
    int axis = 0; float ext = chi[0] - clo[0];
    for (int a = 1; a < 3; ++a) if (chi[a] - clo[a] > ext) { ext = chi[a] - clo[a]; axis = a; }
    rc_axis = axis; rc_cent = b->cent;
    qsort(b->order + first, (size_t)count, sizeof(uint32_t), rc_cmp);
    int64_t half = count / 2;
    int32_t l = rc_build_rec(b, first, half, pad);
    int32_t r = rc_build_rec(b, first + half, count - half, pad);
    b->nodes[me].left = l; b->nodes[me].right = r; b->nodes[me].count = 0;
    return me;
}

void *orc_bvh_build(const float *mesh, const int32_t *ib, int64_t n_tris)
{
    rc_bvh *b = (rc_bvh *)calloc(1, sizeof(rc_bvh));
    b->n_tris = n_tris;
    b->tri = (float *)malloc(sizeof(float) * 9 * (size_t)(n_tris ? n_tris : 1));
    b->cent = (float *)malloc(sizeof(float) * 3 * (size_t)(n_tris ? n_tris : 1));
    b->order = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(n_tris ? n_tris : 1));
    float lo = INFINITY, hi = -INFINITY;
    for (int64_t t = 0; t < n_tris; ++t) {
        for (int k = 0; k < 3; ++k) {
            const float *v = rc_vertex(mesh, ib, t, k);
            for (int a = 0; a < 3; ++a) { b->tri[9 * t + 3 * k + a] = v[a]; lo = fminf(lo, v[a]); hi = fmaxf(hi, v[a]); }
        }
        for (int a = 0; a < 3; ++a) b->cent[3 * t + a] = (b->tri[9 * t + a] + b->tri[9 * t + 3 + a] + b->tri[9 * t + 6 + a]) * (1.0f / 3.0f);
        b->order[t] = (uint32_t)t;
    }
    b->nodes = (rc_node *)malloc(sizeof(rc_node) * (size_t)(2 * n_tris + 1));
    b->n_nodes = 0;
    float pad = n_tris ? (hi - lo) * 1.52587890625e-5f : 0.0f; /* 2^-16 of the extent: generous, CPU side only */
    if (n_tris) rc_build_rec(b, 0, n_tris, pad);
    return b;
}

void orc_bvh_free(void *h)
{
    rc_bvh *b = (rc_bvh *)h;
    if (!b) return;
    free(b->nodes); free(b->order); free(b->tri); free(b->cent); free(b);
}

static inline int rc_slab(const rc_node *n, const float *o, const float *inv, float tmax)
{
    float tn = 0.0f, tf = tmax;
    for (int a = 0; a < 3; ++a) {
        float t0 = (n->lo[a] - o[a]) * inv[a], t1 = (n->hi[a] - o[a]) * inv[a];
        tn = fmaxf(tn, fminf(t0, t1)); /* fminf/fmaxf drop NaN (0*inf): conservative */
        tf = fminf(tf, fmaxf(t0, t1));
    }
    return tn <= tf * 1.0000038f;
}

void orc_bvh_raycast(const void *h, const float *rays, int64_t n_rays, float *out_t, uint32_t *out_id, float *out_u, float *out_v)
{
    const rc_bvh *b = (const rc_bvh *)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (int64_t r = 0; r < n_rays; ++r) {
        const float *o = rays + 8 * r, *d = rays + 8 * r + 4;
        float inv[3] = {1.0f / d[0], 1.0f / d[1], 1.0f / d[2]};
        uint64_t best = ~0ull; float bt = INFINITY, bu = 0, bv = 0;
        int32_t stack[128]; int sp = 0;
        if (b->n_tris) stack[sp++] = 0;
        while (sp) {
            const rc_node *n = &b->nodes[stack[--sp]];
            if (!rc_slab(n, o, inv, bt)) continue;
            if (n->left < 0) {
                for (int32_t i = n->first; i < n->first + n->count; ++i) {
                    uint32_t id = b->order[i]; const float *t = b->tri + 9 * (size_t)id;
                    float tt, u, v;
                    if (rc_moller_trumbore(o, d, t, t + 3, t + 6, &tt, &u, &v)) {
                        uint64_t key = ((uint64_t)rc_bits(tt) << 32) | id;
                        if (key < best) { best = key; bt = tt; bu = u; bv = v; }
                    }
                }
            } else { stack[sp++] = n->left; stack[sp++] = n->right; }
        }
        if (best == ~0ull) { out_t[r] = INFINITY; out_id[r] = ORC_MISS_ID; out_u[r] = 0; out_v[r] = 0; }
        else { out_t[r] = bt; out_id[r] = (uint32_t)best; out_u[r] = bu; out_v[r] = bv; }
    }
}

/* Primary rays for the pixel rect [x0,x0+w) x [y0,y0+h) of a W x H image.
 * cam = {origin[3], U[3], V[3], Wd[3]} (model space):  dir = (U*sx + V*sy) + Wd,
 * sx = (col+0.5)*(2/W) - 1,  sy = 1 - (row+0.5)*(2/H)   (inverse of Dehomogenize, _raster.py:126-129;
 * pixel centres as in _raster.py:298-299). */
void orc_primary_rays(const float *cam, int W, int H, int x0, int y0, int w, int h, float *rays)
{
    const float two_over_w = 2.0f / (float)W, two_over_h = 2.0f / (float)H;
#pragma omp parallel for schedule(static)
    for (int j = 0; j < h; ++j)
        for (int i = 0; i < w; ++i) {
            float sx = ((float)(x0 + i) + 0.5f) * two_over_w - 1.0f;
            float sy = 1.0f - ((float)(y0 + j) + 0.5f) * two_over_h;
            float *r = rays + 8 * ((size_t)j * w + i);
            for (int a = 0; a < 3; ++a) { r[a] = cam[a]; r[4 + a] = (cam[3 + a] * sx + cam[6 + a] * sy) + cam[9 + a]; }
            r[3] = 0; r[7] = 0;
        }
}

static inline uint8_t rc_unorm8(float c)
{
    float v = c * 255.0f;
    if (!(v > 0.0f)) return 0;
    if (v > 255.0f) v = 255.0f;
    return (uint8_t)nearbyintf(v);
}

/* Shade hits into BGRA8.  mode 8: (d,d,d,1), d = max(0.2, N.l) per vertex (lesson08:42), blended
 * d0*(1-u-v) + d1*u + d2*v.  mode 9: tex(C)*L with L = 0.2+max(0,N.l), C = P.xy*2 (lesson09:73,85,93-94).
 * Misses write (0,0,0,0) like clear() (_core.py:386-388). */
void orc_shade_hits(int mode, const float *mesh, const int32_t *ib, const uint32_t *id, const float *u, const float *v,
                    int64_t n, const float *tex, int tex_w, int tex_h, uint8_t *bgra)
{
    const float nn = 0.57735026918962576f;
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < n; ++r) {
        uint8_t *px = bgra + 4 * r;
        if (id[r] == ORC_MISS_ID) { px[0] = px[1] = px[2] = px[3] = 0; continue; }
        float dv[3], cx[3], cy[3];
        for (int k = 0; k < 3; ++k) {
            const float *vt = rc_vertex(mesh, ib, id[r], k);
            float dt = (vt[4] * nn + vt[5] * nn) + vt[6] * nn;
            dv[k] = mode == 8 ? fmaxf(0.2f, dt) : 0.2f + fmaxf(0.0f, dt);
            cx[k] = vt[0] * 2.0f; cy[k] = vt[1] * 2.0f;
        }
        float w0 = 1.0f - u[r] - v[r];
        float d = dv[0] * w0 + dv[1] * u[r] + dv[2] * v[r];
        float c[3] = {d, d, d};
        if (mode == 9) {
            float fx = cx[0] * w0 + cx[1] * u[r] + cx[2] * v[r];
            float fy = cy[0] * w0 + cy[1] * u[r] + cy[2] * v[r];
            float wy = fmodf(fmodf(fy, 1.0f) + 1.0f, 1.0f), wx = fmodf(fmodf(fx, 1.0f) + 1.0f, 1.0f);
            int32_t row = rc_f2i(wy * (float)tex_h), col = rc_f2i(wx * (float)tex_w);
            if (row < 0) row = 0; if (row > tex_h - 1) row = tex_h - 1;
            if (col < 0) col = 0; if (col > tex_w - 1) col = tex_w - 1;
            const float *t = tex + 4 * ((size_t)row * tex_w + col);
            c[0] = t[0] * d; c[1] = t[1] * d; c[2] = t[2] * d;
        }
        px[0] = rc_unorm8(c[2]); px[1] = rc_unorm8(c[1]); px[2] = rc_unorm8(c[0]); px[3] = 255;
    }
}
