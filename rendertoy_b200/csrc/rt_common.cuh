// rt_common.cuh -- shared helpers for librendertoy_b200.so (sm_100a only).
//
// Numerics contract: the whole library is compiled with -fmad=false, default (IEEE) division and sqrt,
// no fast-math.  Every float expression is therefore evaluated exactly as written, left to right, in
// binary32 -- the convention the parity oracle (oracle/*.c) fixes for the reference's OpenCL kernels.
// Where contraction cannot change a result that is compared, code may call __fmaf_rn explicitly.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/rendertoy_b200.h"

void rt_set_error(const char *fmt, ...);

#define RT_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            rt_set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return RT_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define RT_REQUIRE(cond, msg)                                               \
    do {                                                                    \
        if (!(cond)) {                                                      \
            rt_set_error("%s:%d: invalid argument: %s", __FILE__, __LINE__, msg); \
            return RT_ERR_INVALID;                                          \
        }                                                                   \
    } while (0)

int rt_sm_count();

// handle returned by rt_texture_create
struct rt_texture {
    cudaTextureObject_t obj;
    int w, h;
};

// Pixel ownership of an image-space partition ("bin only into owned tiles"; the reference is
// single-device, rendering/_core.py:10-11).  A rank owns the inclusive pixel rect [x0, x1] x [y0, y1] and, inside it, row
// stripes: rows are grouped from y = 0 into stripes of `rows` rows, stripe s is owned iff s % mod == rem (mod = 1: the
// plain rect -- a scissor).  C ABI form: 7 ints {x0, y0, x1, y1, rows, mod, rem}, NULL = the whole frame.
struct RtOwner {
    int x0, y0, x1, y1;
    unsigned rows, mod, rem;
};
__host__ __device__ __forceinline__ bool rt_owns_row(const RtOwner &o, int y)
{
    return y >= o.y0 && y <= o.y1 && (o.mod == 1u || ((unsigned)y / o.rows) % o.mod == o.rem);
}
__host__ __device__ __forceinline__ bool rt_owns(const RtOwner &o, int x, int y) { return x >= o.x0 && x <= o.x1 && rt_owns_row(o, y); }
// any owned row in [ya, yb]?
__host__ __device__ __forceinline__ bool rt_owns_any_row(const RtOwner &o, int ya, int yb)
{
    ya = ya > o.y0 ? ya : o.y0;
    yb = yb < o.y1 ? yb : o.y1;
    if (ya > yb) return false;
    if (o.mod == 1u) return true;
    const unsigned sa = (unsigned)ya / o.rows, sb = (unsigned)yb / o.rows;
    if (sb - sa + 1u >= o.mod) return true;
    for (unsigned s = sa; s <= sb; ++s)
        if (s % o.mod == o.rem) return true;
    return false;
}
// owner7 -> RtOwner clipped to the frame; returns 0 (and the whole frame) for NULL, 1 for a partition, -1 for bad values
static inline int rt_owner_parse(const int32_t *owner7, int width, int height, RtOwner *o)
{
    o->x0 = 0; o->y0 = 0; o->x1 = width - 1; o->y1 = height - 1; o->rows = 1u; o->mod = 1u; o->rem = 0u;
    if (!owner7) return 0;
    if (owner7[4] < 1 || owner7[5] < 1 || owner7[6] < 0 || owner7[6] >= owner7[5]) return -1;
    o->x0 = owner7[0] > 0 ? owner7[0] : 0; o->y0 = owner7[1] > 0 ? owner7[1] : 0;
    o->x1 = owner7[2] < width - 1 ? owner7[2] : width - 1; o->y1 = owner7[3] < height - 1 ? owner7[3] : height - 1;
    o->rows = (unsigned)owner7[4]; o->mod = (unsigned)owner7[5]; o->rem = (unsigned)owner7[6];
    return 1;
}

// prefetch.global.L1 (LEVEL 1) / .L2 (LEVEL 2) of the line holding p; nothing when compiled for the host (tools/cuda_emu)
template <int LEVEL> __device__ __forceinline__ void rt_prefetch(const void *p)
{
#ifdef __CUDACC__
    if (LEVEL == 1) asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
    else asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// normalize((float3)(1,1,1)).x, correctly rounded (lesson08:42)
#define RT_INV_SQRT3 0.57735026918962576f

// write_imagef to CL_UNORM_INT8: saturate, x255, round to nearest even; NaN -> 0
__device__ __forceinline__ uint32_t rt_unorm8(float c)
{
    float v = c * 255.0f;
    v = fminf(fmaxf(v, 0.0f), 255.0f); // fmaxf(NaN, 0) = 0
    return (uint32_t)__float2int_rn(v);
}
// CL_BGRA byte order: B, G, R, A
__device__ __forceinline__ uint32_t rt_pack_bgra(float r, float g, float b, float a)
{
    return rt_unorm8(b) | (rt_unorm8(g) << 8) | (rt_unorm8(r) << 16) | (rt_unorm8(a) << 24);
}

// _core.py:94-96  wrap_coord(c) = fmod(fmod(c, 1) + 1, 1)
__device__ __forceinline__ float rt_wrap(float c) { return fmodf(fmodf(c, 1.0f) + 1.0f, 1.0f); }

// sample2D on a point-sampled float4 texture object over linear memory; texel index computed with the
// reference's arithmetic so selection is bit-identical, clamped to the texture.
__device__ __forceinline__ float4 rt_sample2d(cudaTextureObject_t tex, int tw, int th, float cx, float cy)
{
    int row = (int)(rt_wrap(cy) * (float)th);
    int col = (int)(rt_wrap(cx) * (float)tw);
    row = min(max(row, 0), th - 1);
    col = min(max(col, 0), tw - 1);
    return tex1Dfetch<float4>(tex, row * tw + col);
}
