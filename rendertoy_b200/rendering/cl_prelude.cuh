// cl_prelude.cuh -- device prelude that lets the OpenCL C 1.x subset used by RenderToy scripts compile as CUDA C++
// under NVRTC (prepended to every run-time program by rendertoy_b200/rendering/_dsl.py).
//
// Covers what the reference's tutorials and its own prelude use: vector types with swizzles,
// "(floatN)(...)" literals (rewritten to cl_mk_floatN by _dsl.py), vector arithmetic / relational operators + any(),
// dot / normalize / as_uint / atomics / get_global_id, write_only image2d_t with get_image_dim / write_imagef, the
// float4x4 helpers of rendering/_core.py:41-88 (transpose, rotation, translate, scale, mul) and sample2D (:94-96).
// Compiled with --fmad=false: float expressions evaluate as written.

typedef unsigned int uint;
typedef unsigned long long ulong;   // OpenCL ulong is 64-bit
typedef unsigned char uchar;

struct clf2; struct clf3; struct clf4; struct clf8; struct clf16; struct cli2; struct cli3; struct cli4;

template <typename V, typename T, int STORE, int... I>
struct ClSwz { // view of selected lanes of a parent vector (union member overlaying the parent's storage)
    T s[STORE];
    __device__ operator V() const { V r; int k = 0; ((r.s[k++] = s[I]), ...); return r; }
    __device__ ClSwz &operator=(const V &o) { int k = 0; ((s[I] = o.s[k++]), ...); return *this; }
    __device__ ClSwz &operator+=(const V &o) { int k = 0; ((s[I] = s[I] + o.s[k++]), ...); return *this; }
    __device__ ClSwz &operator-=(const V &o) { int k = 0; ((s[I] = s[I] - o.s[k++]), ...); return *this; }
    __device__ ClSwz &operator*=(const V &o) { int k = 0; ((s[I] = s[I] * o.s[k++]), ...); return *this; }
    __device__ ClSwz &operator/=(const V &o) { int k = 0; ((s[I] = s[I] / o.s[k++]), ...); return *this; }
    __device__ ClSwz &operator*=(T f) { ((s[I] = s[I] * f), ...); return *this; }
    __device__ ClSwz &operator/=(T f) { ((s[I] = s[I] / f), ...); return *this; }
    __device__ ClSwz &operator+=(T f) { ((s[I] = s[I] + f), ...); return *this; }
    __device__ ClSwz &operator-=(T f) { ((s[I] = s[I] - f), ...); return *this; }
};

struct __align__(8) clf2 { union { float s[2]; struct { float x, y; }; }; };
struct __align__(8) cli2 { union { int s[2]; struct { int x, y; }; }; };
struct __align__(16) cli3 { union { int s[4]; struct { int x, y, z; }; }; };
struct __align__(16) cli4 { union { int s[4]; struct { int x, y, z, w; }; }; };
// uintN: CUDA's built-in uint3 is 12 bytes / 4-byte aligned, OpenCL's is 16 / 16 like the host dtype (rendering/_core.py:110)
struct __align__(8) clu2 { union { unsigned s[2]; struct { unsigned x, y; }; }; };
struct __align__(16) clu3 { union { unsigned s[4]; struct { unsigned x, y, z; }; }; };
struct __align__(16) clu4 { union { unsigned s[4]; struct { unsigned x, y, z, w; }; }; };
struct __align__(16) clf3 {
    union { float s[4]; struct { float x, y, z; }; ClSwz<clf2, float, 4, 0, 1> xy; };
};
struct __align__(16) clf4 {
    union { float s[4]; struct { float x, y, z, w; }; ClSwz<clf2, float, 4, 0, 1> xy; ClSwz<clf3, float, 4, 0, 1, 2> xyz; };
};
struct __align__(32) clf8 {
    union {
        float s[8];
        ClSwz<clf4, float, 8, 0, 2, 4, 6> even; ClSwz<clf4, float, 8, 1, 3, 5, 7> odd;
        ClSwz<clf4, float, 8, 0, 1, 2, 3> lo;   ClSwz<clf4, float, 8, 4, 5, 6, 7> hi;
    };
};
struct __align__(64) clf16 {
    union {
        float s[16];
        ClSwz<clf8, float, 16, 0, 2, 4, 6, 8, 10, 12, 14> even; ClSwz<clf8, float, 16, 1, 3, 5, 7, 9, 11, 13, 15> odd;
        ClSwz<clf8, float, 16, 0, 1, 2, 3, 4, 5, 6, 7> lo;      ClSwz<clf8, float, 16, 8, 9, 10, 11, 12, 13, 14, 15> hi;
        ClSwz<clf4, float, 16, 0, 4, 8, 12> even_even_; ClSwz<clf4, float, 16, 1, 5, 9, 13> odd_even_;   // m.even.even etc.,
        ClSwz<clf4, float, 16, 2, 6, 10, 14> even_odd_; ClSwz<clf4, float, 16, 3, 7, 11, 15> odd_odd_;   // renamed by _dsl.py
    };
};

template <typename V> struct ClLanes;
template <> struct ClLanes<clf2> { static const int n = 2; typedef float T; typedef cli2 I; };
template <> struct ClLanes<clf3> { static const int n = 3; typedef float T; typedef cli3 I; };
template <> struct ClLanes<clf4> { static const int n = 4; typedef float T; typedef cli4 I; };
template <> struct ClLanes<cli2> { static const int n = 2; typedef int T; typedef cli2 I; };

#define CL_VEC_OPS(V, T, N)                                                                                                  \
    __device__ inline V operator+(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b.s[i]; return r; } \
    __device__ inline V operator-(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b.s[i]; return r; } \
    __device__ inline V operator*(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b.s[i]; return r; } \
    __device__ inline V operator/(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b.s[i]; return r; } \
    __device__ inline V operator+(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b; return r; }             \
    __device__ inline V operator-(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b; return r; }             \
    __device__ inline V operator*(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b; return r; }             \
    __device__ inline V operator/(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b; return r; }             \
    __device__ inline V operator+(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a + b.s[i]; return r; }             \
    __device__ inline V operator-(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a - b.s[i]; return r; }             \
    __device__ inline V operator*(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a * b.s[i]; return r; }             \
    __device__ inline V operator-(const V &a) { V r; for (int i = 0; i < N; ++i) r.s[i] = -a.s[i]; return r; }                     \
    __device__ inline V &operator+=(V &a, const V &b) { a = a + b; return a; }                                                     \
    __device__ inline V &operator-=(V &a, const V &b) { a = a - b; return a; }                                                     \
    __device__ inline V &operator*=(V &a, const V &b) { a = a * b; return a; }                                                     \
    __device__ inline V &operator/=(V &a, const V &b) { a = a / b; return a; }                                                     \
    __device__ inline V &operator*=(V &a, T b) { a = a * b; return a; }                                                            \
    __device__ inline V &operator/=(V &a, T b) { a = a / b; return a; }                                                            \
    __device__ inline V &operator+=(V &a, T b) { a = a + b; return a; }                                                            \
    __device__ inline V &operator-=(V &a, T b) { a = a - b; return a; }
CL_VEC_OPS(clf2, float, 2)
CL_VEC_OPS(clf3, float, 3)
CL_VEC_OPS(clf4, float, 4)
CL_VEC_OPS(cli2, int, 2)

// float vectors with non-float scalars (int, long, double ...): OpenCL converts the scalar to the element type
template <typename S> struct ClIsScalar { static const bool v = false; };
#define CL_SCALAR(S) template <> struct ClIsScalar<S> { static const bool v = true; };
CL_SCALAR(int) CL_SCALAR(unsigned) CL_SCALAR(long) CL_SCALAR(unsigned long) CL_SCALAR(long long) CL_SCALAR(unsigned long long)
CL_SCALAR(short) CL_SCALAR(unsigned short) CL_SCALAR(char) CL_SCALAR(unsigned char) CL_SCALAR(double)
template <bool B, typename T> struct ClEnableIf { };
template <typename T> struct ClEnableIf<true, T> { typedef T type; };
#define CL_SCALAR_MIX(V)                                                                                                                  \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator*(const V &a, S b) { return a * (float)b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator*(S a, const V &b) { return (float)a * b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator/(const V &a, S b) { return a / (float)b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator+(const V &a, S b) { return a + (float)b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator+(S a, const V &b) { return (float)a + b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator-(const V &a, S b) { return a - (float)b; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, V>::type operator-(S a, const V &b) { return (float)a - b; }
CL_SCALAR_MIX(clf2)
CL_SCALAR_MIX(clf3)
CL_SCALAR_MIX(clf4)

// relational operators: OpenCL yields -1 (all bits) per true lane; any() tests the sign bits
#define CL_REL(V, OPNAME, OP)                                                                                                       \
    __device__ inline typename ClLanes<V>::I OPNAME(const V &a, const V &b) { typename ClLanes<V>::I r; for (int i = 0; i < ClLanes<V>::n; ++i) r.s[i] = (a.s[i] OP b.s[i]) ? -1 : 0; return r; } \
    __device__ inline typename ClLanes<V>::I OPNAME(const V &a, float b) { typename ClLanes<V>::I r; for (int i = 0; i < ClLanes<V>::n; ++i) r.s[i] = (a.s[i] OP b) ? -1 : 0; return r; } \
    template <typename S> __device__ inline typename ClEnableIf<ClIsScalar<S>::v, typename ClLanes<V>::I>::type OPNAME(const V &a, S b) { return OPNAME(a, (float)b); }
#define CL_REL_ALL(V) CL_REL(V, operator<, <) CL_REL(V, operator>, >) CL_REL(V, operator<=, <=) CL_REL(V, operator>=, >=) CL_REL(V, operator==, ==)
CL_REL_ALL(clf2)
CL_REL_ALL(clf3)
CL_REL_ALL(clf4)
__device__ inline int any(const cli2 &v) { return (v.x | v.y) < 0; }
__device__ inline int any(const cli3 &v) { return (v.x | v.y | v.z) < 0; }
__device__ inline int any(const cli4 &v) { return (v.x | v.y | v.z | v.w) < 0; }

// a swizzle on either side of a binary operator behaves like the vector it selects
#define CL_SWZ_BIN(OP)                                                                                                                    \
    template <typename V, typename T, int S, int... I, typename R> __device__ inline auto operator OP(const ClSwz<V, T, S, I...> &a, const R &b) -> decltype(V(a) OP b) { return V(a) OP b; } \
    template <typename V, typename T, int S, int... I> __device__ inline auto operator OP(const V &a, const ClSwz<V, T, S, I...> &b) -> decltype(a OP V(b)) { return a OP V(b); }
CL_SWZ_BIN(+) CL_SWZ_BIN(-) CL_SWZ_BIN(*) CL_SWZ_BIN(/) CL_SWZ_BIN(<) CL_SWZ_BIN(>) CL_SWZ_BIN(<=) CL_SWZ_BIN(>=)

// "(floatN)(...)" literals
__device__ inline clf2 cl_mk_float2(float a, float b) { clf2 r; r.x = a; r.y = b; return r; }
__device__ inline clf2 cl_mk_float2(float a) { return cl_mk_float2(a, a); }
__device__ inline cli2 cl_mk_int2(int a, int b) { cli2 r; r.x = a; r.y = b; return r; }
__device__ inline clf3 cl_mk_float3(float a, float b, float c) { clf3 r; r.x = a; r.y = b; r.z = c; r.s[3] = 0.0f; return r; }
__device__ inline clf3 cl_mk_float3(float a) { return cl_mk_float3(a, a, a); }
__device__ inline clf3 cl_mk_float3(const clf2 &v, float c) { return cl_mk_float3(v.x, v.y, c); }
__device__ inline clf4 cl_mk_float4(float a, float b, float c, float d) { clf4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
__device__ inline clf4 cl_mk_float4(float a) { return cl_mk_float4(a, a, a, a); }
__device__ inline clf4 cl_mk_float4(const clf3 &v, float d) { return cl_mk_float4(v.x, v.y, v.z, d); }
__device__ inline clf4 cl_mk_float4(const clf2 &v, float c, float d) { return cl_mk_float4(v.x, v.y, c, d); }
__device__ inline clf16 cl_mk_float16(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8, float a9,
                                      float a10, float a11, float a12, float a13, float a14, float a15)
{
    clf16 r;
    const float v[16] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};
    for (int i = 0; i < 16; ++i) r.s[i] = v[i];
    return r;
}

// builtins (min/max/fmod/sin/cos/exp/pow/sqrt on scalars come from CUDA's own overloads)
__device__ inline float dot(const clf2 &a, const clf2 &b) { return a.x * b.x + a.y * b.y; }
__device__ inline float dot(const clf3 &a, const clf3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__device__ inline float dot(const clf4 &a, const clf4 &b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
__device__ inline clf3 normalize(const clf3 &v) { float l = sqrtf(dot(v, v)); return v / l; }
__device__ inline clf4 normalize(const clf4 &v) { float l = sqrtf(dot(v, v)); return v / l; }
__device__ inline clf2 normalize(const clf2 &v) { float l = sqrtf(dot(v, v)); return v / l; }
__device__ inline float length(const clf3 &v) { return sqrtf(dot(v, v)); }
__device__ inline clf3 cross(const clf3 &a, const clf3 &b) { return cl_mk_float3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ inline uint as_uint(float f) { return __float_as_uint(f); }
__device__ inline float as_float(uint u) { return __uint_as_float(u); }
__device__ inline int atomic_add(int *p, int v) { return atomicAdd(p, v); }
__device__ inline uint atomic_min(uint *p, uint v) { return atomicMin(p, v); }
__device__ inline int get_global_id(int) { return (int)(blockIdx.x * blockDim.x + threadIdx.x); }

// images: linear device memory, row-major; BGRA8 targets store B,G,R,A (CL_BGRA / CL_UNORM_INT8, rendering/_core.py:340)
struct ClImage { void *data; int width, height, components, is_unorm8_bgra; };
typedef ClImage image2d_t;
__device__ inline cli2 get_image_dim(const ClImage &im) { return cl_mk_int2(im.width, im.height); }
__device__ inline uint cl_unorm8(float c) { float v = fminf(fmaxf(c * 255.0f, 0.0f), 255.0f); return (uint)__float2int_rn(v); }
__device__ inline void write_imagef(const ClImage &im, cli2 p, clf4 c)
{
    if (p.x < 0 || p.y < 0 || p.x >= im.width || p.y >= im.height) return;
    const size_t i = (size_t)p.y * im.width + p.x;
    if (im.is_unorm8_bgra) ((uint *)im.data)[i] = cl_unorm8(c.z) | (cl_unorm8(c.y) << 8) | (cl_unorm8(c.x) << 16) | (cl_unorm8(c.w) << 24);
    else for (int k = 0; k < im.components; ++k) ((float *)im.data)[i * im.components + k] = c.s[k];
}

// float4x4 helpers with the semantics of rendering/_core.py:41-88 (row-vector convention, row-major float16)
typedef clf16 cl_float4x4;
__device__ inline clf4 mul(const clf4 &v, const clf16 &m)
{
    clf4 r;
    for (int j = 0; j < 4; ++j) r.s[j] = ((v.x * m.s[j] + v.y * m.s[4 + j]) + v.z * m.s[8 + j]) + v.w * m.s[12 + j];
    return r;
}
__device__ inline clf16 transpose(const clf16 &m) { clf16 t; for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) t.s[4 * i + j] = m.s[4 * j + i]; return t; }
__device__ inline clf16 rotation(float angle, clf3 axis)
{
    const float c = cosf(angle), s = sinf(angle), x = axis.x, y = axis.y, z = axis.z;
    return cl_mk_float16(x * x * (1 - c) + c, y * x * (1 - c) + z * s, z * x * (1 - c) - y * s, 0,
                         x * y * (1 - c) - z * s, y * y * (1 - c) + c, z * y * (1 - c) + x * s, 0,
                         x * z * (1 - c) + y * s, y * z * (1 - c) - x * s, z * z * (1 - c) + c, 0,
                         0, 0, 0, 1);
}
__device__ inline clf16 translate(clf3 v) { return cl_mk_float16(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, v.x, v.y, v.z, 1); }
__device__ inline clf16 scale(clf3 v) { return cl_mk_float16(v.x, 0, 0, 0, 0, v.y, 0, 0, 0, 0, v.z, 0, 0, 0, 0, 1); }

// sample2D (rendering/_core.py:94-96): nearest texel with repeat wrap from the texture pool; index clamped to the texture
#define wrap_coord(c) (fmodf(fmodf((c), 1.0f) + 1.0f, 1.0f))
template <typename TEX>
__device__ inline clf4 cl_sample2D(unsigned long long pool, const TEX &t, const clf2 &c)
{
    int row = (int)(wrap_coord(c.y) * t.height), col = (int)(wrap_coord(c.x) * t.width);
    row = min(max(row, 0), t.height - 1); col = min(max(col, 0), t.width - 1);
    return ((const clf4 *)(pool + (unsigned long long)t.offset))[row * t.width + col];
}

// sample2D_linear: the bilinear sampler docs/09_texture_mapping.md:69-70 asks for ("would be another function, like
// sample2D_linear"); the reference never got it.  Same repeat wrap as sample2D, texel centres at (i + 0.5) / size (so a
// coordinate that sample2D maps to the middle of a texel returns exactly that texel), neighbours wrap around the edges,
// weights and blends in float32 as written.
template <typename TEX>
__device__ inline clf4 cl_sample2D_linear(unsigned long long pool, const TEX &t, const clf2 &c)
{
    const float x = wrap_coord(c.x) * t.width - 0.5f, y = wrap_coord(c.y) * t.height - 0.5f;
    const float fx = floorf(x), fy = floorf(y);
    const float ax = x - fx, ay = y - fy;
    int c0 = (int)fx, r0 = (int)fy;
    c0 = c0 < 0 ? c0 + t.width : (c0 >= t.width ? c0 - t.width : c0);
    r0 = r0 < 0 ? r0 + t.height : (r0 >= t.height ? r0 - t.height : r0);
    const int c1 = c0 + 1 >= t.width ? 0 : c0 + 1, r1 = r0 + 1 >= t.height ? 0 : r0 + 1;
    const clf4 *tex = (const clf4 *)(pool + (unsigned long long)t.offset);
    const clf4 t00 = tex[r0 * t.width + c0], t01 = tex[r0 * t.width + c1], t10 = tex[r1 * t.width + c0], t11 = tex[r1 * t.width + c1];
    const clf4 top = t00 * (1.0f - ax) + t01 * ax, bottom = t10 * (1.0f - ax) + t11 * ax;
    return top * (1.0f - ay) + bottom * ay;
}

#define __kernel
#define __global
#define __constant const
#define __local
#define __private
#define read_only
#define write_only
#define float2 clf2
#define float3 clf3
#define float4 clf4
#define float8 clf8
#define float16 clf16
#define float4x4 clf16
#define int2 cli2
#define int3 cli3
#define int4 cli4
#define uint2 clu2
#define uint3 clu3
#define uint4 clu4
