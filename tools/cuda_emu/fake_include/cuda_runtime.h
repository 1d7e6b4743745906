// tools/cuda_emu -- DEV-TIME TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// A just-enough CUDA execution model for the host: lets the SOURCE of a kernel file (csrc/*.cu, launches rewritten by
// tools/cuda_emu/build.py) run on CPU threads -- one std::thread per CUDA thread, blocks one after the other, __syncthreads and
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
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }
inline int2 make_int2(int x, int y) { return {x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }

struct EmuDim3 { unsigned x = 1, y = 1, z = 1; };
inline thread_local EmuDim3 threadIdx, blockIdx, blockDim, gridDim;

typedef void *cudaStream_t;
typedef unsigned long long cudaTextureObject_t;
enum cudaError_t { cudaSuccess = 0 };
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char *cudaGetErrorString(cudaError_t) { return "emulated"; }
template <typename T> inline T tex1Dfetch(cudaTextureObject_t tex, int i) { return reinterpret_cast<const T *>((uintptr_t)tex)[i]; }

inline thread_local unsigned long long t_emu_loads = 0;     // per-thread __ldg / __ldcg calls, summed into warp_loads / 32 at exit
template <typename T> inline T __ldg(const T *p) { ++t_emu_loads; return *p; }
template <typename T> inline T __ldcg(const T *p) { ++t_emu_loads; return *p; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float2int_rn(float f) { return (int)nearbyintf(f); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
inline int __ffs(unsigned v) { return v ? __builtin_ctz(v) + 1 : 0; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }

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
    std::barrier<> bar{32};
    unsigned long long slot[32];
    std::atomic<unsigned> alive{0xffffffffu};
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
template <typename F> inline void emu_launch(unsigned grid, unsigned block, F body)
{
    if (block % 32 != 0) { fprintf(stderr, "cuda_emu: block size must be a multiple of 32\n"); abort(); }
    for (unsigned b = 0; b < grid; ++b) {
        EmuBlock blk;
        blk.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)block);
        for (unsigned w = 0; w < block / 32; ++w) blk.warps.push_back(std::make_unique<EmuWarp>());
        g_emu_block = &blk;
        std::vector<std::thread> threads;
        threads.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            threads.emplace_back([&, t] {
                threadIdx.x = t; blockIdx.x = b; blockDim.x = block; gridDim.x = grid;
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
