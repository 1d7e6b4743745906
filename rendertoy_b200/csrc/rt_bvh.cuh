// rt_bvh.cuh -- device layout of the LBVH shared by the builder (rt_bvh.cu) and the traversal (rt_raycast.cu).
#pragma once
#include <cuda_runtime.h>

// Inner node, 64 B = four 128-bit loads.  Holds BOTH children's (padded) boxes so one visit decides both.
//   n0 = (c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y)   n1 = (c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y)
//   n2 = (c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z)   n3 = (child0, child1, -, -): >= 0 inner index, < 0 leaf ~slot
struct __align__(16) RtBvhNode { float4 n0, n1, n2; int4 n3; };

// Leaf triangle in sorted order, 48 B: v0.w carries the ORIGINAL triangle id (bits); edges precomputed
// (the same float subtraction Moller-Trumbore starts with, so results stay bit-identical to the oracle).
struct __align__(16) RtBvhTri { float4 v0, e1, e2; };

// rt_bvh_build's `builder` (include/rendertoy_b200.h) and the bound both builders keep: the traversal stacks hold
// one pending sibling per level, 64 entries (Karras: 32 key bits + index tie-break; PLOC: checked after the build).
#ifndef RT_BVH_LBVH
#define RT_BVH_LBVH 0
#define RT_BVH_PLOC 1
#endif
#define RT_BVH_MAX_HEIGHT 64
