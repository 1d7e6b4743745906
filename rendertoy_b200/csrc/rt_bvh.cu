// rt_bvh.cu -- GPU LBVH build for Raycaster._build_ads (rendering/_raycaster.py:30-33, a no-op in the reference;
// struct names BVH_AABB / BVH_Triangle at :8-22).
//
//   1. triangle centroids -> scene bounds (ordered-int atomics)
//   2. 30-bit Morton codes (10 bits/axis) of the normalised centroids
//   3. LSD radix sort of (code, triangle id), 4 x 8-bit passes: per-tile histogram -> exclusive scan ->
//      stable scatter (warp match_any ranking); equal codes keep triangle-id order
//   4. Karras 2012 hierarchy: one thread per internal node finds its key range and split
//   5. bottom-up refit with per-node arrival counters; boxes padded so that any hit the float32
//      Moller-Trumbore test reports lies inside every ancestor box
//   6. emit traversal layout: 64 B inner nodes holding BOTH children's boxes (one visit = 4 x 128-bit loads),
//      48 B leaf triangles {v0 | id, e1, e2} in sorted order
//
// All of this is HBM-/latency-bound integer and min/max work: no contraction, no tensor cores.
#include "rt_common.cuh"
#include "rt_bvh.cuh"

namespace {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;

__device__ __forceinline__ int float_to_ordered(float f)
{
    int i = __float_as_int(f);
    return i >= 0 ? i : i ^ 0x7fffffff;
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

struct TriV { float4 a, b, c; };

__device__ __forceinline__ TriV load_tri(const float4 *pos, const int *idx, long long t)
{
    long long i0 = 3 * t, i1 = 3 * t + 1, i2 = 3 * t + 2;
    if (idx) { i0 = idx[i0]; i1 = idx[i1]; i2 = idx[i2]; }
    TriV v;
    v.a = __ldg(pos + i0); v.b = __ldg(pos + i1); v.c = __ldg(pos + i2);
    return v;
}

__device__ __forceinline__ float3 centroid(const TriV &v)
{
    const float third = 1.0f / 3.0f;
    return make_float3((v.a.x + v.b.x + v.c.x) * third, (v.a.y + v.b.y + v.c.y) * third, (v.a.z + v.b.z + v.c.z) * third);
}

// bounds[0..2] = min centroid, [3..5] = max centroid, [6..8] = min vertex, [9..11] = max vertex (ordered ints)
__global__ void bounds_init_kernel(int *bounds)
{
    if (threadIdx.x < 12) bounds[threadIdx.x] = (threadIdx.x % 6) < 3 ? 0x7fffffff : (int)0x80000000;
}

__global__ void bounds_kernel(const float4 *pos, const int *idx, int n, int *bounds)
{
    float lo[6] = {INFINITY, INFINITY, INFINITY, INFINITY, INFINITY, INFINITY};
    float hi[6] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY, -INFINITY};
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
        TriV v = load_tri(pos, idx, t);
        float3 c = centroid(v);
        lo[0] = fminf(lo[0], c.x); lo[1] = fminf(lo[1], c.y); lo[2] = fminf(lo[2], c.z);
        hi[0] = fmaxf(hi[0], c.x); hi[1] = fmaxf(hi[1], c.y); hi[2] = fmaxf(hi[2], c.z);
        lo[3] = fminf(lo[3], fminf(v.a.x, fminf(v.b.x, v.c.x))); hi[3] = fmaxf(hi[3], fmaxf(v.a.x, fmaxf(v.b.x, v.c.x)));
        lo[4] = fminf(lo[4], fminf(v.a.y, fminf(v.b.y, v.c.y))); hi[4] = fmaxf(hi[4], fmaxf(v.a.y, fmaxf(v.b.y, v.c.y)));
        lo[5] = fminf(lo[5], fminf(v.a.z, fminf(v.b.z, v.c.z))); hi[5] = fmaxf(hi[5], fmaxf(v.a.z, fmaxf(v.b.z, v.c.z)));
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) {
            lo[k] = fminf(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], d));
            hi[k] = fmaxf(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], d));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(bounds + k, float_to_ordered(lo[k]));      atomicMax(bounds + 3 + k, float_to_ordered(hi[k]));
            atomicMin(bounds + 6 + k, float_to_ordered(lo[3 + k])); atomicMax(bounds + 9 + k, float_to_ordered(hi[3 + k]));
        }
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void morton_kernel(const float4 *pos, const int *idx, int n, const int *bounds, uint32_t *keys, uint32_t *vals)
{
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    float3 c = centroid(load_tri(pos, idx, t));
    float q[3] = {c.x, c.y, c.z};
    uint32_t code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float lo = ordered_to_float(bounds[k]), hi = ordered_to_float(bounds[3 + k]);
        float ext = hi - lo;
        float u = ext > 0.0f ? (q[k] - lo) / ext : 0.0f;
        int cell = (int)(u * 1024.0f);
        cell = min(max(cell, 0), 1023);
        code |= expand_bits10((uint32_t)cell) << (2 - k);
    }
    keys[t] = code;
    vals[t] = (uint32_t)t;
}

// ---- radix sort ---------------------------------------------------------------------------------------

__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const uint32_t *keys, int n, int shift, uint32_t *counts, int nblocks)
{
    __shared__ unsigned h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int base = blockIdx.x * SORT_TILE;
#pragma unroll
    for (int i = 0; i < SORT_ITEMS; ++i) {
        int g = base + i * SORT_THREADS + threadIdx.x;
        if (g < n) atomicAdd(&h[(keys[g] >> shift) & 255u], 1u);
    }
    __syncthreads();
    counts[threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x]; // digit-major, so one scan orders (digit, tile)
}

__global__ void __launch_bounds__(1024) sort_scan_kernel(uint32_t *counts, int total)
{
    __shared__ unsigned warp_sums[32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned carry = 0;
    for (int base = 0; base < total; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < total ? counts[i] : 0u;
        unsigned incl = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned x = __shfl_up_sync(0xffffffffu, incl, d);
            if (lane >= d) incl += x;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            unsigned w = warp_sums[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                unsigned x = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += x;
            }
            warp_sums[lane] = w;
        }
        __syncthreads();
        const unsigned prefix = wid ? warp_sums[wid - 1] : 0u;
        if (i < total) counts[i] = carry + prefix + incl - v;
        carry += warp_sums[31];
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SORT_THREADS) sort_scatter_kernel(const uint32_t *keys_in, const uint32_t *vals_in, uint32_t *keys_out,
                                                                     uint32_t *vals_out, int n, int shift, const uint32_t *offsets, int nblocks)
{
    __shared__ unsigned cnt[SORT_THREADS / 32][256];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < (SORT_THREADS / 32) * 256; i += SORT_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    // warp `wid` owns 256 consecutive keys of the tile, taken 32 at a time so in-warp order == input order
    const int wbase = blockIdx.x * SORT_TILE + wid * (32 * SORT_ITEMS);
    uint32_t key[SORT_ITEMS], val[SORT_ITEMS];
    unsigned rank[SORT_ITEMS];
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int g = wbase + r * 32 + lane;
        const bool valid = g < n;
        key[r] = valid ? keys_in[g] : 0u;
        val[r] = valid ? vals_in[g] : 0u;
        const unsigned d = valid ? ((key[r] >> shift) & 255u) : 256u;
        const unsigned peers = __match_any_sync(FULL, d);
        const unsigned before = __popc(peers & ((1u << lane) - 1u));
        unsigned prev = 0;
        if (valid) prev = cnt[wid][d];
        __syncwarp();
        if (valid && before == 0) cnt[wid][d] = prev + __popc(peers);
        __syncwarp();
        rank[r] = prev + before;
    }
    __syncthreads();
    { // digit `threadIdx.x`: global offset of this tile, then running base per warp
        const unsigned d = threadIdx.x;
        unsigned run = offsets[d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) {
            unsigned c = cnt[w][d];
            cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ITEMS; ++r) {
        const int g = wbase + r * 32 + lane;
        if (g < n) {
            const unsigned p = cnt[wid][(key[r] >> shift) & 255u] + rank[r];
            keys_out[p] = key[r];
            vals_out[p] = val[r];
        }
    }
}

// ---- Karras hierarchy -------------------------------------------------------------------------------------

__device__ __forceinline__ int delta(const uint32_t *keys, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const uint32_t a = keys[i], b = keys[j];
    return a != b ? __clz(a ^ b) : 32 + __clz((uint32_t)i ^ (uint32_t)j); // index breaks ties of equal codes
}

// children: >= 0 inner node index, < 0 leaf ~slot.  parent[] is indexed inner: i, leaf: (n - 1) + slot.
__global__ void karras_kernel(const uint32_t *keys, int n, int2 *children, int *parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = delta(keys, n, i, i - d);
    int lmax = 2;
    while (delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int left = min(i, j) == gamma ? ~gamma : gamma;
    const int right = max(i, j) == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    children[i] = make_int2(left, right);
    parent[left >= 0 ? left : (n - 1) + ~left] = i;
    parent[right >= 0 ? right : (n - 1) + ~right] = i;
    if (i == 0) parent[0] = -1;
}

// ---- leaves + refit -----------------------------------------------------------------------------------------

struct Box { float3 lo, hi; };

__device__ __forceinline__ Box box_union(const Box &a, const Box &b)
{
    Box r;
    r.lo = make_float3(fminf(a.lo.x, b.lo.x), fminf(a.lo.y, b.lo.y), fminf(a.lo.z, b.lo.z));
    r.hi = make_float3(fmaxf(a.hi.x, b.hi.x), fmaxf(a.hi.y, b.hi.y), fmaxf(a.hi.z, b.hi.z));
    return r;
}

__device__ __forceinline__ Box load_box(const float4 *boxes, int k)
{
    const float4 a = __ldcg(boxes + 2 * k), b = __ldcg(boxes + 2 * k + 1);
    Box r;
    r.lo = make_float3(a.x, a.y, a.z); r.hi = make_float3(b.x, b.y, b.z);
    return r;
}

__device__ __forceinline__ void store_box(float4 *boxes, int k, const Box &b)
{
    __stcg(boxes + 2 * k, make_float4(b.lo.x, b.lo.y, b.lo.z, 0.0f));
    __stcg(boxes + 2 * k + 1, make_float4(b.hi.x, b.hi.y, b.hi.z, 0.0f));
}

// One thread per sorted leaf: write its traversal triangle and padded box, then climb; the second thread to
// reach an inner node merges its children's boxes (box index: inner i, leaf (n-1)+slot) and goes on.
__global__ void leaves_refit_kernel(const float4 *pos, const int *idx, int n, const uint32_t *sorted_ids, const int *bounds,
                                    const int2 *children, const int *parent, int *arrivals, float4 *boxes, RtBvhTri *tris,
                                    RtBvhNode *nodes)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n) return;
    const uint32_t id = sorted_ids[slot];
    const TriV v = load_tri(pos, idx, id);
    RtBvhTri tr;
    tr.v0 = make_float4(v.a.x, v.a.y, v.a.z, __uint_as_float(id));
    tr.e1 = make_float4(v.b.x - v.a.x, v.b.y - v.a.y, v.b.z - v.a.z, 0.0f);
    tr.e2 = make_float4(v.c.x - v.a.x, v.c.y - v.a.y, v.c.z - v.a.z, 0.0f);
    tris[slot] = tr;
    // pad = 2^-17 of the largest scene extent (vertex bounds)
    float ext = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) ext = fmaxf(ext, ordered_to_float(bounds[9 + k]) - ordered_to_float(bounds[6 + k]));
    const float pad = ext * 7.62939453125e-6f;
    Box b;
    b.lo = make_float3(fminf(v.a.x, fminf(v.b.x, v.c.x)) - pad, fminf(v.a.y, fminf(v.b.y, v.c.y)) - pad, fminf(v.a.z, fminf(v.b.z, v.c.z)) - pad);
    b.hi = make_float3(fmaxf(v.a.x, fmaxf(v.b.x, v.c.x)) + pad, fmaxf(v.a.y, fmaxf(v.b.y, v.c.y)) + pad, fmaxf(v.a.z, fmaxf(v.b.z, v.c.z)) + pad);
    store_box(boxes, (n - 1) + slot, b);
    if (n == 1) return;
    __threadfence();
    int node = parent[(n - 1) + slot];
    while (node >= 0) {
        if (atomicAdd(arrivals + node, 1) == 0) return; // first to arrive: the sibling subtree is not done yet
        __threadfence();
        const int2 ch = children[node];
        const Box l = load_box(boxes, ch.x >= 0 ? ch.x : (n - 1) + ~ch.x);
        const Box r = load_box(boxes, ch.y >= 0 ? ch.y : (n - 1) + ~ch.y);
        RtBvhNode nd;
        nd.n0 = make_float4(l.lo.x, l.hi.x, l.lo.y, l.hi.y);
        nd.n1 = make_float4(r.lo.x, r.hi.x, r.lo.y, r.hi.y);
        nd.n2 = make_float4(l.lo.z, l.hi.z, r.lo.z, r.hi.z);
        nd.n3 = make_int4(ch.x, ch.y, 0, 0);
        nodes[node] = nd;
        store_box(boxes, node, box_union(l, r));
        __threadfence();
        node = parent[node];
    }
}

// ---- PLOC: parallel locally-ordered clustering (Meister & Bittner 2018) ------------------------------------------
// Bottom-up agglomeration over the Morton-sorted leaves: every cluster looks PLOC_RADIUS neighbours to either side for
// the partner with the smallest merged surface area; mutual pairs merge into a new inner node, the survivors are
// compacted (order preserved) and the round repeats until one cluster is left.  Measured on dragon100k: SAH cost of
// the inner nodes 45.6 against 52.8 for the Karras tree and 40.8 for a full-sweep SAH build on the CPU (radius 4 .. 64
// moves it by less than 1), 20.5 instead of 21.8 node visits per ray and 3 - 6 % off the cfg4 frame, for about 1.3 ms
// more build time.  No host round trips: the cluster count lives in device memory, every kernel of a round exits when it
// is 1, and the host enqueues a fixed number of rounds -- PLOC rounds first, then rounds that pair clusters 2k / 2k+1
// unconditionally (halving, so any input finishes).  Node ids are handed out downwards from n-2, which leaves the root
// at 0 where the traversal expects it.  Each cluster carries its subtree height; the caller checks it against the
// traversal's stack bound and falls back to the Karras tree (height <= 64 by construction) if it is exceeded.
#ifndef RT_PLOC_RADIUS
#define RT_PLOC_RADIUS 32
#endif
constexpr int PLOC_RADIUS = RT_PLOC_RADIUS;
constexpr int PLOC_TB = 256;
constexpr int PLOC_ROUNDS = 72;   // nearest-neighbour rounds ...
constexpr int PAIR_ROUNDS = 32;   // ... then forced pairing: 2^32 > any cluster count left

struct PlocCtl { int count[2]; int next_id; int height; }; // count[round & 1] = clusters entering that round

__device__ __forceinline__ float half_area(const float4 &lo, const float4 &hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return dx * dy + dy * dz + dz * dx;
}

// one thread per sorted leaf: traversal triangle + padded box as cluster `slot` (box.lo.w = ref bits, box.hi.w = height)
__global__ void ploc_leaves_kernel(const float4 *pos, const int *idx, int n, const uint32_t *sorted_ids, const int *bounds, float4 *cbox,
                                   RtBvhTri *tris, PlocCtl *ctl)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot == 0) { ctl->count[0] = n; ctl->count[1] = n; ctl->next_id = n - 2; ctl->height = 0; }
    if (slot >= n) return;
    const uint32_t id = sorted_ids[slot];
    const TriV v = load_tri(pos, idx, id);
    RtBvhTri tr;
    tr.v0 = make_float4(v.a.x, v.a.y, v.a.z, __uint_as_float(id));
    tr.e1 = make_float4(v.b.x - v.a.x, v.b.y - v.a.y, v.b.z - v.a.z, 0.0f);
    tr.e2 = make_float4(v.c.x - v.a.x, v.c.y - v.a.y, v.c.z - v.a.z, 0.0f);
    tris[slot] = tr;
    float ext = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) ext = fmaxf(ext, ordered_to_float(bounds[9 + k]) - ordered_to_float(bounds[6 + k]));
    const float pad = ext * 7.62939453125e-6f; // 2^-17 of the largest scene extent, as leaves_refit_kernel
    cbox[2 * slot] = make_float4(fminf(v.a.x, fminf(v.b.x, v.c.x)) - pad, fminf(v.a.y, fminf(v.b.y, v.c.y)) - pad,
                                 fminf(v.a.z, fminf(v.b.z, v.c.z)) - pad, __int_as_float(~slot));
    cbox[2 * slot + 1] = make_float4(fmaxf(v.a.x, fmaxf(v.b.x, v.c.x)) + pad, fmaxf(v.a.y, fmaxf(v.b.y, v.c.y)) + pad,
                                     fmaxf(v.a.z, fmaxf(v.b.z, v.c.z)) + pad, __int_as_float(0));
}

// nearest neighbour by merged area within the radius; ties go to the lower index, which guarantees a mutual pair
__global__ void __launch_bounds__(PLOC_TB) ploc_nn_kernel(const float4 *cbox, const PlocCtl *ctl, int round, int *nn)
{
    __shared__ float4 slo[PLOC_TB + 2 * PLOC_RADIUS], shi[PLOC_TB + 2 * PLOC_RADIUS];
    const int c = ctl->count[round & 1];
    const int base = blockIdx.x * PLOC_TB;
    if (c <= 1 || base >= c) return;
    for (int k = threadIdx.x; k < PLOC_TB + 2 * PLOC_RADIUS; k += PLOC_TB) {
        const int g = base - PLOC_RADIUS + k;
        if (g >= 0 && g < c) { slo[k] = cbox[2 * g]; shi[k] = cbox[2 * g + 1]; }
    }
    __syncthreads();
    const int i = base + threadIdx.x;
    if (i >= c) return;
    if (round >= PLOC_ROUNDS) { nn[i] = (i ^ 1) < c ? (i ^ 1) : i; return; } // forced pairing rounds
    const float4 lo = slo[threadIdx.x + PLOC_RADIUS], hi = shi[threadIdx.x + PLOC_RADIUS];
    float best = INFINITY;
    int bj = i;
    for (int k = -PLOC_RADIUS; k <= PLOC_RADIUS; ++k) {
        const int j = i + k;
        if (k == 0 || j < 0 || j >= c) continue;
        const float4 l2 = slo[threadIdx.x + PLOC_RADIUS + k], h2 = shi[threadIdx.x + PLOC_RADIUS + k];
        const float4 ul = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), 0.0f);
        const float4 uh = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), 0.0f);
        const float a = half_area(ul, uh);
        if (a < best || (bj == i && !(a > best))) { best = a; bj = j; } // strict '<' keeps the lowest index among equals; NaN areas still pick someone
    }
    nn[i] = bj;
}

// flags[i]: bit 0 = cluster i survives into the next round (alone, or as the leader of a mutual pair), bit 1 = leader.
// Block totals (survivors, leaders) go to sums[2 * block], sums[2 * block + 1].
__global__ void __launch_bounds__(PLOC_TB) ploc_flag_kernel(const PlocCtl *ctl, int round, const int *nn, int *flags, int *sums)
{
    __shared__ int wsum[2][PLOC_TB / 32];
    const int c = ctl->count[round & 1];
    const int base = blockIdx.x * PLOC_TB;
    if (c <= 1 || base >= c) return;
    const int i = base + threadIdx.x;
    int f = 0;
    if (i < c) {
        const int j = nn[i];
        const bool mutual = j != i && nn[j] == i;
        f = !mutual ? 1 : (i < j ? 3 : 0);
        flags[i] = f;
    }
    const unsigned keep = __ballot_sync(0xffffffffu, f & 1), lead = __ballot_sync(0xffffffffu, f & 2);
    if ((threadIdx.x & 31) == 0) { wsum[0][threadIdx.x >> 5] = __popc(keep); wsum[1][threadIdx.x >> 5] = __popc(lead); }
    __syncthreads();
    if (threadIdx.x < 2) {
        int t = 0;
        for (int w = 0; w < PLOC_TB / 32; ++w) t += wsum[threadIdx.x][w];
        sums[2 * blockIdx.x + threadIdx.x] = t;
    }
}

// exclusive scan of the per-block totals (one block; at most 2^30 / 256 = 4M entries, in practice a few thousand),
// then the next round's cluster count and node-id watermark
__global__ void __launch_bounds__(1024) ploc_scan_kernel(PlocCtl *ctl, int round, int *sums)
{
    __shared__ int wsum[2][32];
    __shared__ int carry[2];
    const int c = ctl->count[round & 1];
    if (c <= 1) { // finished: keep both slots at the final count, so that the remaining rounds stay no-ops
        if (threadIdx.x == 0) ctl->count[(round + 1) & 1] = c;
        return;
    }
    const int nb = (c + PLOC_TB - 1) / PLOC_TB;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (threadIdx.x < 2) carry[threadIdx.x] = 0;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int b = base + threadIdx.x;
        int v[2] = {b < nb ? sums[2 * b] : 0, b < nb ? sums[2 * b + 1] : 0};
        int incl[2] = {v[0], v[1]};
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int x0 = __shfl_up_sync(0xffffffffu, incl[0], d), x1 = __shfl_up_sync(0xffffffffu, incl[1], d);
            if (lane >= d) { incl[0] += x0; incl[1] += x1; }
        }
        if (lane == 31) { wsum[0][wid] = incl[0]; wsum[1][wid] = incl[1]; }
        __syncthreads();
        if (wid < 2) {
            int w = wsum[wid][lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int x = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += x;
            }
            wsum[wid][lane] = w;
        }
        __syncthreads();
        if (b < nb) {
            sums[2 * b] = carry[0] + (wid ? wsum[0][wid - 1] : 0) + incl[0] - v[0];
            sums[2 * b + 1] = carry[1] + (wid ? wsum[1][wid - 1] : 0) + incl[1] - v[1];
        }
        __syncthreads();
        if (threadIdx.x < 2) carry[threadIdx.x] += wsum[threadIdx.x][31];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        ctl->count[(round + 1) & 1] = carry[0];
        sums[2 * nb] = ctl->next_id;   // this round's leaders take ids next_id, next_id - 1, ... (read by ploc_merge_kernel)
        ctl->next_id -= carry[1];
    }
}

// survivors move to their compacted position; leaders first become an inner node over the two partners
__global__ void __launch_bounds__(PLOC_TB) ploc_merge_kernel(const float4 *cin, float4 *cout, PlocCtl *ctl, int round, const int *nn, const int *flags,
                                                             const int *sums, RtBvhNode *nodes)
{
    __shared__ int wsum[2][PLOC_TB / 32];
    const int c = ctl->count[round & 1];
    const int base = blockIdx.x * PLOC_TB;
    if (c <= 1 || base >= c) return;
    const int nb = (c + PLOC_TB - 1) / PLOC_TB;
    const int i = base + threadIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int f = i < c ? flags[i] : 0;
    const unsigned keep = __ballot_sync(0xffffffffu, f & 1), lead = __ballot_sync(0xffffffffu, f & 2);
    if (lane == 0) { wsum[0][wid] = __popc(keep); wsum[1][wid] = __popc(lead); }
    __syncthreads();
    int pos = sums[2 * blockIdx.x], rank = sums[2 * blockIdx.x + 1];
    for (int w = 0; w < wid; ++w) { pos += wsum[0][w]; rank += wsum[1][w]; }
    pos += __popc(keep & ((1u << lane) - 1u));
    rank += __popc(lead & ((1u << lane) - 1u));
    if (!(f & 1)) return;
    float4 lo = cin[2 * i], hi = cin[2 * i + 1];
    if (f & 2) {
        const int j = nn[i];
        const float4 l2 = cin[2 * j], h2 = cin[2 * j + 1];
        const int id = sums[2 * nb] - rank;
        RtBvhNode nd;
        nd.n0 = make_float4(lo.x, hi.x, lo.y, hi.y);
        nd.n1 = make_float4(l2.x, h2.x, l2.y, h2.y);
        nd.n2 = make_float4(lo.z, hi.z, l2.z, h2.z);
        nd.n3 = make_int4(__float_as_int(lo.w), __float_as_int(l2.w), 0, 0);
        nodes[id] = nd;
        const int height = 1 + max(__float_as_int(hi.w), __float_as_int(h2.w));
        lo = make_float4(fminf(lo.x, l2.x), fminf(lo.y, l2.y), fminf(lo.z, l2.z), __int_as_float(id));
        hi = make_float4(fmaxf(hi.x, h2.x), fmaxf(hi.y, h2.y), fmaxf(hi.z, h2.z), __int_as_float(height));
        if (id == 0) ctl->height = height; // the root
    }
    cout[2 * pos] = lo;
    cout[2 * pos + 1] = hi;
}

// n == 1: a single inner node whose both children are the only leaf's box / an empty box
__global__ void single_leaf_root_kernel(const float4 *boxes, RtBvhNode *nodes)
{
    const Box l = load_box(boxes, 0);
    RtBvhNode nd;
    nd.n0 = make_float4(l.lo.x, l.hi.x, l.lo.y, l.hi.y);
    nd.n1 = make_float4(INFINITY, -INFINITY, INFINITY, -INFINITY);
    nd.n2 = make_float4(l.lo.z, l.hi.z, INFINITY, -INFINITY);
    nd.n3 = make_int4(~0, ~0, 0, 0);
    nodes[0] = nd;
}

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct ScratchLayout {
    size_t keys0, keys1, vals0, vals1, counts, bounds, children, parent, arrivals, boxes, total;
    int nblocks;
};

ScratchLayout scratch_layout(int64_t n)
{
    ScratchLayout L;
    const size_t nn = (size_t)(n > 0 ? n : 1);
    L.nblocks = (int)((nn + SORT_TILE - 1) / SORT_TILE);
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    L.keys0 = take(nn * 4); L.keys1 = take(nn * 4); L.vals0 = take(nn * 4); L.vals1 = take(nn * 4);
    L.counts = take((size_t)L.nblocks * 256 * 4);
    L.bounds = take(12 * 4);
    L.children = take(nn * 8);
    L.parent = take(2 * nn * 4);
    L.arrivals = take(nn * 4);
    L.boxes = take(2 * nn * 32);
    L.total = off;
    return L;
}

} // namespace

extern "C" {

int64_t rt_bvh_node_bytes(int64_t n_triangles) { return (int64_t)sizeof(RtBvhNode) * (n_triangles > 1 ? n_triangles - 1 : 1); }
int64_t rt_bvh_tri_bytes(int64_t n_triangles) { return (int64_t)sizeof(RtBvhTri) * (n_triangles > 0 ? n_triangles : 1); }
int64_t rt_bvh_scratch_bytes(int64_t n_triangles) { return (int64_t)scratch_layout(n_triangles).total; }

int rt_bvh_build(const void *d_pos4, const int32_t *d_indices, int64_t n_triangles, void *d_nodes, void *d_tris, void *d_scratch,
                 int builder, void *stream)
{
    RT_REQUIRE(builder == RT_BVH_LBVH || builder == RT_BVH_PLOC, "builder must be RT_BVH_LBVH or RT_BVH_PLOC");
    RT_REQUIRE(n_triangles >= 1 && n_triangles < (1ll << 30), "triangle count must be in [1, 2^30)");
    RT_REQUIRE(d_pos4 && d_nodes && d_tris && d_scratch, "buffers");
    RT_REQUIRE((((uintptr_t)d_pos4 | (uintptr_t)d_nodes | (uintptr_t)d_tris | (uintptr_t)d_scratch) & 15) == 0, "16-byte alignment");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = (int)n_triangles;
    const ScratchLayout L = scratch_layout(n);
    char *s = (char *)d_scratch;
    uint32_t *keys[2] = {(uint32_t *)(s + L.keys0), (uint32_t *)(s + L.keys1)};
    uint32_t *vals[2] = {(uint32_t *)(s + L.vals0), (uint32_t *)(s + L.vals1)};
    uint32_t *counts = (uint32_t *)(s + L.counts);
    int *bounds = (int *)(s + L.bounds);
    int2 *children = (int2 *)(s + L.children);
    int *parent = (int *)(s + L.parent);
    int *arrivals = (int *)(s + L.arrivals);
    float4 *boxes = (float4 *)(s + L.boxes);
    const float4 *pos = (const float4 *)d_pos4;
    const int tb = 256, gb = (n + tb - 1) / tb;

    bounds_init_kernel<<<1, 32, 0, st>>>(bounds);
    int red_blocks = gb < rt_sm_count() * 8 ? gb : rt_sm_count() * 8;
    bounds_kernel<<<red_blocks, tb, 0, st>>>(pos, d_indices, n, bounds);
    morton_kernel<<<gb, tb, 0, st>>>(pos, d_indices, n, bounds, keys[0], vals[0]);
    RT_CUDA(cudaGetLastError());
    int cur = 0;
    for (int shift = 0; shift < 32; shift += 8) {
        sort_hist_kernel<<<L.nblocks, SORT_THREADS, 0, st>>>(keys[cur], n, shift, counts, L.nblocks);
        sort_scan_kernel<<<1, 1024, 0, st>>>(counts, L.nblocks * 256);
        sort_scatter_kernel<<<L.nblocks, SORT_THREADS, 0, st>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], n, shift, counts, L.nblocks);
        cur ^= 1;
    }
    RT_CUDA(cudaGetLastError());
    if (builder == RT_BVH_PLOC && n >= 3) {
        // scratch reuse: clusters (box + ref + height, ping-pong) <- boxes, nearest neighbours <- arrivals, flags <- parent,
        // block sums <- counts, control block <- children; the sorted keys / ids stay intact for the fallback below
        float4 *cbox[2] = {boxes, boxes + 2 * (size_t)n};
        PlocCtl *ctl = (PlocCtl *)children;
        int *nn = arrivals, *flags = parent, *sums = (int *)counts;
        ploc_leaves_kernel<<<gb, tb, 0, st>>>(pos, d_indices, n, vals[cur], bounds, cbox[0], (RtBvhTri *)d_tris, ctl);
        // a round at least halves nothing and at most halves everything; the grid shrinks with the guaranteed bound
        // (forced rounds halve), PLOC rounds keep the full grid (blocks past the live count return at once)
        const int pb = (n + PLOC_TB - 1) / PLOC_TB;
        for (int round = 0; round < PLOC_ROUNDS + PAIR_ROUNDS; ++round) {
            ploc_nn_kernel<<<pb, PLOC_TB, 0, st>>>(cbox[round & 1], ctl, round, nn);
            ploc_flag_kernel<<<pb, PLOC_TB, 0, st>>>(ctl, round, nn, flags, sums);
            ploc_scan_kernel<<<1, 1024, 0, st>>>(ctl, round, sums);
            ploc_merge_kernel<<<pb, PLOC_TB, 0, st>>>(cbox[round & 1], cbox[(round + 1) & 1], ctl, round, nn, flags, sums, (RtBvhNode *)d_nodes);
        }
        RT_CUDA(cudaGetLastError());
        PlocCtl h;
        RT_CUDA(cudaMemcpyAsync(&h, ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
        RT_CUDA(cudaStreamSynchronize(st));
        if (h.count[0] == 1 && h.count[1] == 1 && h.next_id == -1 && h.height <= RT_BVH_MAX_HEIGHT) return RT_OK;
        // not reached in practice: a tree deeper than the traversal stack (or an unfinished one) -- rebuild as LBVH
    }
    RT_CUDA(cudaMemsetAsync(arrivals, 0, (size_t)n * 4, st));
    if (n > 1) karras_kernel<<<(n - 1 + tb - 1) / tb, tb, 0, st>>>(keys[cur], n, children, parent);
    leaves_refit_kernel<<<gb, tb, 0, st>>>(pos, d_indices, n, vals[cur], bounds, children, parent, arrivals, boxes, (RtBvhTri *)d_tris,
                                           (RtBvhNode *)d_nodes);
    if (n == 1) single_leaf_root_kernel<<<1, 1, 0, st>>>(boxes, (RtBvhNode *)d_nodes);
    RT_CUDA(cudaGetLastError());
    return RT_OK;
}

} // extern "C"
