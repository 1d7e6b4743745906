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
