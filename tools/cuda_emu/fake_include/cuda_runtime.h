// tools/cuda_emu -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// A just-enough CUDA execution model for the host: lets the SOURCE of a kernel file (csrc/*.cu, launches rewritten by
// tools/cuda_emu/emu_build.py) run on CPU threads -- one std::thread per CUDA thread, blocks one after the other, __syncthreads and
// the warp collectives as barriers -- so that kernel LOGIC (indexing, shared-memory protocols, vote / reduce sequences,
// conservativeness of tests) can be checked against the oracle without a GPU.  It says nothing about performance, memory
// ordering subtleties or anything the hardware does differently from "32 lanes arriving at every collective".
#pragma once
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) int2 { int x, y; };
struct alignas(8) uint2 { unsigned x, y; };
struct alignas(16) ulonglong2 { unsigned long long x, y; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct alignas(8) float3 { float x, y, z; };
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return {x, y, z, w}; }
inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return {x, y}; }
inline float3 make_float3(float x, float y, float z) { return {x, y, z}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }

struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned x_, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
    dim3(int x_, int y_ = 1, int z_ = 1) : x((unsigned)x_), y((unsigned)y_), z((unsigned)z_) {}
};
typedef dim3 EmuDim3;
inline thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

typedef void *cudaStream_t;
typedef unsigned long long cudaTextureObject_t;
enum cudaError_t { cudaSuccess = 0 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t = nullptr) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, int, cudaStream_t = nullptr) { memcpy(d, s, n); return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
// "device" memory is host memory here; copies are synchronous; IPC handles carry the pointer itself
inline cudaError_t cudaMalloc(void **p, size_t n) { *p = malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
struct cudaIpcMemHandle_t { char reserved[64]; };
enum { cudaIpcMemLazyEnablePeerAccess = 1 };
inline cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) { memset(h, 0, sizeof *h); memcpy(h->reserved, &p, sizeof p); return cudaSuccess; }
inline cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) { memcpy(p, h.reserved, sizeof *p); return cudaSuccess; }
inline cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void *d, size_t dp, const void *s, size_t sp, size_t w, size_t h, int, cudaStream_t = nullptr)
{
    for (size_t r = 0; r < h; ++r) memcpy((char *)d + r * dp, (const char *)s + r * sp, w);
    return cudaSuccess;
}
struct cudaPitchedPtr { void *ptr; size_t pitch, xsize, ysize; };
struct cudaExtent { size_t width, height, depth; };
struct cudaPos { size_t x, y, z; };
struct cudaMemcpy3DParms { void *srcArray; cudaPos srcPos; cudaPitchedPtr srcPtr; void *dstArray; cudaPos dstPos; cudaPitchedPtr dstPtr; cudaExtent extent; int kind; };
inline cudaPitchedPtr make_cudaPitchedPtr(void *p, size_t pitch, size_t xs, size_t ys) { return cudaPitchedPtr{p, pitch, xs, ys}; }
inline cudaExtent make_cudaExtent(size_t w, size_t h, size_t d) { return cudaExtent{w, h, d}; }
inline cudaError_t cudaMemcpy3DAsync(const cudaMemcpy3DParms *p, cudaStream_t = nullptr)
{   // linear memory on both sides: a slice is ysize rows of pitch bytes
    for (size_t z = 0; z < p->extent.depth; ++z)
        for (size_t y = 0; y < p->extent.height; ++y)
            memcpy((char *)p->dstPtr.ptr + ((z + p->dstPos.z) * p->dstPtr.ysize + y + p->dstPos.y) * p->dstPtr.pitch + p->dstPos.x,
                   (const char *)p->srcPtr.ptr + ((z + p->srcPos.z) * p->srcPtr.ysize + y + p->srcPos.y) * p->srcPtr.pitch + p->srcPos.x, p->extent.width);
    return cudaSuccess;
}
// linear-memory, point-sampled texture objects only: the "object" is the texel pointer
template <typename T> inline T tex1Dfetch(cudaTextureObject_t tex, int i) { return reinterpret_cast<const T *>((uintptr_t)tex)[i]; }
enum { cudaResourceTypeLinear = 2, cudaFilterModePoint = 0, cudaReadModeElementType = 0 };
struct cudaChannelFormatDesc { int x, y, z, w, f; };
template <typename T> inline cudaChannelFormatDesc cudaCreateChannelDesc() { return {(int)sizeof(T) * 2, 0, 0, 0, 0}; }
struct cudaResourceDesc { int resType; struct { struct { void *devPtr; cudaChannelFormatDesc desc; size_t sizeInBytes; } linear; } res; };
struct cudaTextureDesc { int filterMode, readMode, normalizedCoords; };
inline cudaError_t cudaCreateTextureObject(cudaTextureObject_t *obj, const cudaResourceDesc *rd, const cudaTextureDesc *, const void *)
{
    *obj = (cudaTextureObject_t)(uintptr_t)rd->res.linear.devPtr;
    return cudaSuccess;
}
inline cudaError_t cudaDestroyTextureObject(cudaTextureObject_t) { return cudaSuccess; }

inline thread_local unsigned long long t_emu_loads = 0;     // per-thread __ldg / __ldcg calls, summed into warp_loads / 32 at exit
template <typename T> inline T __ldg(const T *p) { ++t_emu_loads; return *p; }
template <typename T> inline T __ldcg(const T *p) { ++t_emu_loads; return *p; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
inline int __float2int_ru(float f) { return (int)ceilf(f); }
inline int __float2int_rd(float f) { return (int)floorf(f); }
inline float __frcp_rn(float f) { return 1.0f / f; }
template <typename T> inline T __ldcs(const T *p) { return *p; }
inline int atomicSub(int *p, int v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicSub(unsigned *p, unsigned v) { return __atomic_fetch_sub(p, v, __ATOMIC_SEQ_CST); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }

inline float __fdividef(float a, float b) { return a / b; }     // the GPU's is approximate: callers must not depend on its rounding
inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; ++i) r |= ((v >> i) & 1u) << (31 - i); return r; }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __fns(unsigned mask, unsigned base, int offset)
{   // position of the offset-th set bit of mask at or above base (offset > 0), 0xffffffff if there is none
    int seen = 0;
    if (offset > 0) { for (unsigned i = base; i < 32; ++i) if ((mask >> i) & 1u) { if (++seen == offset) return i; } }
    else if (offset < 0) { for (int i = (int)base; i >= 0; --i) if ((mask >> i) & 1u) { if (++seen == -offset) return (unsigned)i; } }
    else if ((mask >> base) & 1u) return base;
    return 0xffffffffu;
}
inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v)
{
    unsigned long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMin(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int atomicMin(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMax(int *p, int v)
{
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned atomicMax(unsigned *p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
template <typename T> inline void __stcg(T *p, T v) { *p = v; }
template <typename T> inline void __stcs(T *p, T v) { *p = v; }
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaMemcpyDefault = 4 };
inline unsigned atomicCAS(unsigned *p, unsigned expected, unsigned desired)
{
    __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
    return expected; // the value found at *p (== the caller's `expected` iff the swap happened)
}
inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }

// warp-level event counters (one count per warp, not per lane): a hardware-independent proxy for how much a kernel variant
// executes -- a node visit of the packet walk is 3 wide loads + 2 votes, a pop 1 vote, a leaf visit 3 wide loads
struct EmuCounts { std::atomic<unsigned long long> votes{0}, reduces{0}, shuffles{0}, warp_loads{0}, block_syncs{0}; };
inline EmuCounts g_emu_counts;
extern "C" __attribute__((used, visibility("default"))) void emu_counts_read(unsigned long long *out5, int reset)
{
    out5[0] = g_emu_counts.votes; out5[1] = g_emu_counts.reduces; out5[2] = g_emu_counts.shuffles; out5[3] = g_emu_counts.warp_loads;
    out5[4] = g_emu_counts.block_syncs;
    if (reset) { g_emu_counts.votes = 0; g_emu_counts.reduces = 0; g_emu_counts.shuffles = 0; g_emu_counts.warp_loads = 0; g_emu_counts.block_syncs = 0; }
}

// ---- block / warp state of the block that is currently running ------------------------------------------------------
struct EmuWarp {
    std::barrier<> bar;
    unsigned long long slot[32];
    std::atomic<unsigned> alive;
    explicit EmuWarp(unsigned lanes) : bar((std::ptrdiff_t)lanes), alive(lanes >= 32 ? 0xffffffffu : ((1u << lanes) - 1u)) {}
};
struct EmuBlock {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<std::unique_ptr<EmuWarp>> warps;
};
inline EmuBlock *g_emu_block = nullptr;
inline EmuWarp &emu_warp() { return *g_emu_block->warps[threadIdx.x >> 5]; }

inline void __syncthreads()
{
    if (threadIdx.x == 0) ++g_emu_counts.block_syncs;
    g_emu_block->bar->arrive_and_wait();
}
inline void __syncwarp(unsigned = 0xffffffffu) { emu_warp().bar.arrive_and_wait(); }
// block-wide OR of a predicate (every thread of the block must call it, as on the device)
inline int __syncthreads_or(int pred)
{
    static std::atomic<int> acc[2];            // blocks run one after the other; two phases so a fast thread cannot clear too early
    static std::atomic<unsigned> phase{0};
    const unsigned ph = phase.load() & 1u;
    if (pred) acc[ph].store(1);
    g_emu_block->bar->arrive_and_wait();
    const int r = acc[ph].load();
    g_emu_block->bar->arrive_and_wait();
    if (threadIdx.x == 0) { acc[ph].store(0); phase.fetch_add(1); }
    g_emu_block->bar->arrive_and_wait();
    return r;
}

// every collective: publish, barrier, read all live lanes, barrier (so nobody overwrites a slot somebody still reads)
template <typename F> inline auto emu_collective(unsigned long long mine, F combine)
{
    EmuWarp &w = emu_warp();
    w.slot[threadIdx.x & 31] = mine;
    w.bar.arrive_and_wait();
    auto r = combine(w.slot, w.alive.load());
    w.bar.arrive_and_wait();
    return r;
}
inline unsigned __ballot_sync(unsigned, bool pred)
{
    if ((threadIdx.x & 31) == 0) ++g_emu_counts.votes;
    return emu_collective(pred ? 1ull : 0ull, [](const unsigned long long *s, unsigned alive) {
        unsigned m = 0;
        for (int i = 0; i < 32; ++i)
            if (((alive >> i) & 1u) && s[i]) m |= 1u << i;
        return m;
    });
}
inline int __shfl_sync(unsigned, int v, int src)
{
    if ((threadIdx.x & 31) == 0) ++g_emu_counts.shuffles;
    return emu_collective((unsigned long long)(unsigned)v, [src](const unsigned long long *s, unsigned) { return (int)(unsigned)s[src & 31]; });
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0u; }
inline bool __all_sync(unsigned m, bool pred) { return __ballot_sync(m, !pred) == 0u; }
inline int __shfl_up_sync(unsigned, int v, unsigned delta)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) ++g_emu_counts.shuffles;
    return emu_collective((unsigned long long)(unsigned)v, [lane, delta](const unsigned long long *s, unsigned) {
        return (int)(unsigned)s[lane >= (int)delta ? lane - (int)delta : lane];
    });
}
inline int __shfl_xor_sync(unsigned, int v, int lane_mask)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) ++g_emu_counts.shuffles;
    return emu_collective((unsigned long long)(unsigned)v, [lane, lane_mask](const unsigned long long *s, unsigned) { return (int)(unsigned)s[(lane ^ lane_mask) & 31]; });
}
inline unsigned __shfl_xor_sync(unsigned m, unsigned v, int lane_mask) { return (unsigned)__shfl_xor_sync(m, (int)v, lane_mask); }
inline float __shfl_xor_sync(unsigned m, float v, int lane_mask) { return __int_as_float(__shfl_xor_sync(m, __float_as_int(v), lane_mask)); }
inline unsigned __match_any_sync(unsigned, unsigned v)
{
    const int lane = threadIdx.x & 31;
    if (lane == 0) ++g_emu_counts.votes;
    return emu_collective((unsigned long long)v, [lane](const unsigned long long *s, unsigned alive) {
        unsigned m = 0;
        for (int i = 0; i < 32; ++i)
            if (((alive >> i) & 1u) && s[i] == s[lane]) m |= 1u << i;
        return m;
    });
}
inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
    if ((threadIdx.x & 31) == 0) ++g_emu_counts.reduces;
    return emu_collective((unsigned long long)v, [](const unsigned long long *s, unsigned alive) {
        unsigned m = 0;
        for (int i = 0; i < 32; ++i)
            if ((alive >> i) & 1u) m |= (unsigned)s[i];
        return m;
    });
}
inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
    if ((threadIdx.x & 31) == 0) ++g_emu_counts.reduces;
    return emu_collective((unsigned long long)v, [](const unsigned long long *s, unsigned alive) {
        unsigned m = 0xffffffffu;
        for (int i = 0; i < 32; ++i)
            if ((alive >> i) & 1u) m = (unsigned)s[i] < m ? (unsigned)s[i] : m;
        return m;
    });
}

// kernel<<<grid, block>>>(args)  ->  emu_launch(grid, block, [&] { kernel(args); })
template <typename F> inline void emu_launch(dim3 grid3, dim3 block3, F body)
{
    const unsigned grid = grid3.x * grid3.y * grid3.z, block = block3.x;
    if (block == 0 || block3.y != 1 || block3.z != 1) { fprintf(stderr, "cuda_emu: 1-D blocks only\n"); abort(); }
    for (unsigned b = 0; b < grid; ++b) {
        EmuBlock blk;
        blk.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        for (unsigned w = 0; w * 32 < block; ++w) blk.warps.push_back(std::make_unique<EmuWarp>(block - w * 32 < 32 ? block - w * 32 : 32u));
        g_emu_block = &blk;
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t] {
                threadIdx.x = t; blockDim.x = block; gridDim = grid3;
                blockIdx.x = b % grid3.x; blockIdx.y = (b / grid3.x) % grid3.y; blockIdx.z = b / (grid3.x * grid3.y);
                t_emu_loads = 0;
                body();
                g_emu_counts.warp_loads += t_emu_loads;  // thread-level; readers divide by 32 for a warp-level lower bound
                EmuWarp &w = *blk.warps[t >> 5];          // an exited thread no longer takes part in anything
                w.alive.fetch_and(~(1u << (t & 31)));
                w.bar.arrive_and_drop();
                blk.bar->arrive_and_drop();
            });
        for (auto &th : threads) th.join();
        g_emu_block = nullptr;
    }
}
