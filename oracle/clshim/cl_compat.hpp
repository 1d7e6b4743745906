// oracle/clshim/cl_compat.hpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Just enough of OpenCL C 1.x, in C++17, to compile the reference's OWN kernel text (rendering/_core.py:38-98
// prelude, the kernels rendering/_raster.py generates, and the tutorial shaders) for the host CPU, so the
// unmodified reference Python can run here without pyopencl/an OpenCL ICD.  oracle/clshim/translate.py does
// the few textual rewrites C++ needs ("(float4)(a,b,c,d)" -> "make_float4(a,b,c,d)", address-space keywords).
//
// Float semantics: strict binary32, no contraction (compiled -ffp-contract=off, no fast-math); dot() sums left
// to right; normalize() divides by sqrtf(dot) -- the conventions oracle/raster_oracle.c states.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

typedef unsigned int uint;
typedef unsigned long ulong;
typedef unsigned char uchar;

#define __kernel
#define __global
#define __constant static const
#define __local
#define __private
#define read_only
#define write_only

// ---- vector types with OpenCL swizzles ------------------------------------------------------------------
struct float2; struct float3; struct float4; struct float8; struct float16; struct int2;

// View of selected lanes of a parent vector (a union member overlaying the parent's storage of STORE lanes):
// converts to / assigns from the vector type V made of those lanes.
template <typename V, typename T, int STORE, int... I>
struct SwzImpl {
    T s[STORE];
    operator V() const { V r; int k = 0; ((r.s[k++] = s[I]), ...); return r; }
    SwzImpl &operator=(const V &o) { int k = 0; ((s[I] = o.s[k++]), ...); return *this; }
    SwzImpl &operator+=(const V &o) { int k = 0; ((s[I] = s[I] + o.s[k++]), ...); return *this; }
    SwzImpl &operator-=(const V &o) { int k = 0; ((s[I] = s[I] - o.s[k++]), ...); return *this; }
    SwzImpl &operator*=(const V &o) { int k = 0; ((s[I] = s[I] * o.s[k++]), ...); return *this; }
    SwzImpl &operator/=(const V &o) { int k = 0; ((s[I] = s[I] / o.s[k++]), ...); return *this; }
    SwzImpl &operator*=(T f) { ((s[I] = s[I] * f), ...); return *this; }
    SwzImpl &operator/=(T f) { ((s[I] = s[I] / f), ...); return *this; }
    SwzImpl &operator+=(T f) { ((s[I] = s[I] + f), ...); return *this; }
    SwzImpl &operator-=(T f) { ((s[I] = s[I] - f), ...); return *this; }
};

#define VEC_OPS(V, T, N)                                                                                       \
    inline V operator+(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b.s[i]; return r; } \
    inline V operator-(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b.s[i]; return r; } \
    inline V operator*(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b.s[i]; return r; } \
    inline V operator/(const V &a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b.s[i]; return r; } \
    inline V operator+(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] + b; return r; }             \
    inline V operator-(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] - b; return r; }             \
    inline V operator*(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] * b; return r; }             \
    inline V operator/(const V &a, T b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a.s[i] / b; return r; }             \
    inline V operator+(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a + b.s[i]; return r; }             \
    inline V operator-(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a - b.s[i]; return r; }             \
    inline V operator*(T a, const V &b) { V r; for (int i = 0; i < N; ++i) r.s[i] = a * b.s[i]; return r; }             \
    inline V operator-(const V &a) { V r; for (int i = 0; i < N; ++i) r.s[i] = -a.s[i]; return r; }                     \
    inline V &operator+=(V &a, const V &b) { a = a + b; return a; }                                                     \
    inline V &operator-=(V &a, const V &b) { a = a - b; return a; }                                                     \
    inline V &operator*=(V &a, const V &b) { a = a * b; return a; }                                                     \
    inline V &operator/=(V &a, const V &b) { a = a / b; return a; }                                                     \
    inline V &operator*=(V &a, T b) { a = a * b; return a; }                                                            \
    inline V &operator/=(V &a, T b) { a = a / b; return a; }

struct alignas(8) float2 {
    union { float s[2]; struct { float x, y; }; };
};
struct alignas(8) int2 {
    union { int s[2]; struct { int x, y; }; };
};
struct alignas(16) float3 {
    union {
        float s[4];
        struct { float x, y, z; };
        SwzImpl<float2, float, 4, 0, 1> xy;
    };
};
struct alignas(16) float4 {
    union {
        float s[4];
        struct { float x, y, z, w; };
        SwzImpl<float2, float, 4, 0, 1> xy;
        SwzImpl<float3, float, 4, 0, 1, 2> xyz;
    };
};
struct alignas(32) float8 {
    union {
        float s[8];
        SwzImpl<float4, float, 8, 0, 2, 4, 6> even;
        SwzImpl<float4, float, 8, 1, 3, 5, 7> odd;
        SwzImpl<float4, float, 8, 0, 1, 2, 3> lo;
        SwzImpl<float4, float, 8, 4, 5, 6, 7> hi;
    };
};
struct alignas(64) float16 {
    union {
        float s[16];
        SwzImpl<float8, float, 16, 0, 2, 4, 6, 8, 10, 12, 14> even;
        SwzImpl<float8, float, 16, 1, 3, 5, 7, 9, 11, 13, 15> odd;
        SwzImpl<float8, float, 16, 0, 1, 2, 3, 4, 5, 6, 7> lo;
        SwzImpl<float8, float, 16, 8, 9, 10, 11, 12, 13, 14, 15> hi;
        // two-level swizzles of the prelude's mul() (m.even.even ...), which translate.py renames
        SwzImpl<float4, float, 16, 0, 4, 8, 12> even_even_;
        SwzImpl<float4, float, 16, 1, 5, 9, 13> odd_even_;
        SwzImpl<float4, float, 16, 2, 6, 10, 14> even_odd_;
        SwzImpl<float4, float, 16, 3, 7, 11, 15> odd_odd_;
    };
};
typedef float16 float4x4;

VEC_OPS(float2, float, 2)
VEC_OPS(float3, float, 3)
VEC_OPS(float4, float, 4)
VEC_OPS(int2, int, 2)

// mixed swizzle arithmetic the kernels use: (swizzle OP scalar/vector) yields the vector type
template <typename V, typename T, int S, int... I> inline V operator*(const SwzImpl<V, T, S, I...> &a, T b) { return V(a) * b; }
template <typename V, typename T, int S, int... I> inline V operator*(const SwzImpl<V, T, S, I...> &a, int b) { return V(a) * (T)b; }
template <typename V, typename T, int S, int... I> inline V operator*(const SwzImpl<V, T, S, I...> &a, const V &b) { return V(a) * b; }
template <typename V, typename T, int S, int... I> inline V operator+(const SwzImpl<V, T, S, I...> &a, const V &b) { return V(a) + b; }
template <typename V, typename T, int S, int... I> inline V operator-(const SwzImpl<V, T, S, I...> &a, const V &b) { return V(a) - b; }
inline float2 operator*(const float2 &a, int b) { return a * (float)b; }
inline float3 operator*(const float3 &a, int b) { return a * (float)b; }

// relational operators (OpenCL: -1 per true lane) + any(), as used by the tutorials' clip tests
struct alignas(16) int3 { union { int s[4]; struct { int x, y, z; }; }; };
#define REL_OPS(OPNAME, OP)                                                                                                   \
    inline int3 OPNAME(const float3 &a, const float3 &b) { int3 r; for (int i = 0; i < 3; ++i) r.s[i] = (a.s[i] OP b.s[i]) ? -1 : 0; r.s[3] = 0; return r; } \
    inline int3 OPNAME(const float3 &a, double b) { int3 r; for (int i = 0; i < 3; ++i) r.s[i] = (a.s[i] OP (float)b) ? -1 : 0; r.s[3] = 0; return r; }     \
    template <int S, int... I> inline int3 OPNAME(const SwzImpl<float3, float, S, I...> &a, const float3 &b) { return OPNAME(float3(a), b); }                \
    template <int S, int... I> inline int3 OPNAME(const SwzImpl<float3, float, S, I...> &a, double b) { return OPNAME(float3(a), b); }
REL_OPS(operator<, <) REL_OPS(operator>, >) REL_OPS(operator<=, <=) REL_OPS(operator>=, >=)
inline int any(const int3 &v) { return (v.x | v.y | v.z) < 0; }
inline float3 operator+(const float3 &a, double b) { return a + (float)b; }

// ---- "(floatN)(...)" constructors (translate.py rewrites the cast syntax to these) -----------------------
inline float2 make_float2(float a, float b) { float2 r; r.x = a; r.y = b; return r; }
inline float2 make_float2(float a) { return make_float2(a, a); }
inline int2 make_int2(int a, int b) { int2 r; r.x = a; r.y = b; return r; }
inline float3 make_float3(float a, float b, float c) { float3 r; r.x = a; r.y = b; r.z = c; r.s[3] = 0.0f; return r; }
inline float3 make_float3(float a) { return make_float3(a, a, a); }
inline float4 make_float4(float a, float b, float c, float d) { float4 r; r.x = a; r.y = b; r.z = c; r.w = d; return r; }
inline float4 make_float4(float a) { return make_float4(a, a, a, a); }
inline float4 make_float4(const float3 &v, float d) { return make_float4(v.x, v.y, v.z, d); }
inline float4 make_float4(const float2 &v, float c, float d) { return make_float4(v.x, v.y, c, d); }
inline float16 make_float16(float a0, float a1, float a2, float a3, float a4, float a5, float a6, float a7, float a8, float a9, float a10,
                            float a11, float a12, float a13, float a14, float a15)
{
    float16 r;
    const float v[16] = {a0, a1, a2, a3, a4, a5, a6, a7, a8, a9, a10, a11, a12, a13, a14, a15};
    for (int i = 0; i < 16; ++i) r.s[i] = v[i];
    return r;
}
#define make_float4x4 make_float16

// ---- builtins ------------------------------------------------------------------------------------------
inline float dot(const float2 &a, const float2 &b) { return a.x * b.x + a.y * b.y; }
inline float dot(const float3 &a, const float3 &b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
inline float dot(const float4 &a, const float4 &b) { return ((a.x * b.x + a.y * b.y) + a.z * b.z) + a.w * b.w; }
template <typename T, int S, int... I> inline float dot(const float4 &a, const SwzImpl<float4, T, S, I...> &b) { return dot(a, float4(b)); }
inline float3 normalize(const float3 &v) { float l = sqrtf(dot(v, v)); return v / l; }
inline float4 normalize(const float4 &v) { float l = sqrtf(dot(v, v)); return v / l; }
inline float max(float a, float b) { return fmaxf(a, b); }
inline float min(float a, float b) { return fminf(a, b); }
inline int max(int a, int b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline float max(double a, float b) { return fmaxf((float)a, b); }
inline float max(float a, double b) { return fmaxf(a, (float)b); }
inline float fmod(float a, float b) { return fmodf(a, b); }
inline float cos(float a) { return cosf(a); }
inline float sin(float a) { return sinf(a); }
inline float sqrt(float a) { return sqrtf(a); }
inline float exp(float a) { return expf(a); }
inline float pow(float a, float b) { return powf(a, b); }
inline uint as_uint(float f) { uint u; memcpy(&u, &f, 4); return u; }
inline float as_float(uint u) { float f; memcpy(&f, &u, 4); return f; }

inline int atomic_add(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline uint atomic_min(uint *p, uint v)
{
    uint cur = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (v < cur && !__atomic_compare_exchange_n(p, &cur, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) { }
    return cur;
}

// (int) casts in the kernels: the reference ran on GPUs, where float->int saturates and NaN gives 0
struct SatInt {
    static int cvt(float f)
    {
        if (f != f) return 0;
        if (f >= 2147483648.0f) return INT32_MAX;
        if (f <= -2147483648.0f) return INT32_MIN;
        return (int)f;
    }
    static int cvt(double f) { return cvt((float)f); }
    static int cvt(int i) { return i; }
    static int cvt(long i) { return (int)i; }
    static int cvt(unsigned i) { return (int)i; }
};

static thread_local long __cl_gid = 0;
inline int get_global_id(int) { return (int)__cl_gid; }

// ---- images ------------------------------------------------------------------------------------------------
struct ClImage {
    void *data;
    int width, height;
    int components; // channels per pixel
    int is_unorm8_bgra;
};
typedef ClImage *image2d_t;
inline int2 get_image_dim(image2d_t im) { return make_int2(im->width, im->height); }

inline unsigned char cl_unorm8(float c) // convert_uchar_sat_rte(c * 255.0f)
{
    float v = c * 255.0f;
    if (!(v > 0.0f)) return 0;
    if (v > 255.0f) v = 255.0f;
    return (unsigned char)nearbyintf(v);
}

inline void write_imagef(image2d_t im, int2 p, float4 c)
{
    if (p.x < 0 || p.y < 0 || p.x >= im->width || p.y >= im->height) return; // out-of-range writes are dropped by hardware
    if (im->is_unorm8_bgra) {
        unsigned char *px = (unsigned char *)im->data + 4 * ((size_t)p.y * im->width + p.x);
        px[0] = cl_unorm8(c.z); px[1] = cl_unorm8(c.y); px[2] = cl_unorm8(c.x); px[3] = cl_unorm8(c.w);
    } else {
        float *px = (float *)im->data + (size_t)im->components * ((size_t)p.y * im->width + p.x);
        for (int i = 0; i < im->components; ++i) px[i] = c.s[i];
    }
}
