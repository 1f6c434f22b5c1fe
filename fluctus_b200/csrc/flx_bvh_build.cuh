// flx_bvh_build.cuh -- GPU construction of the acceleration structure the traversal stages consume.
//
// The reference builds its hierarchy on the CPU (src/bvh.cpp:205-407 full-sweep SAH, src/sbvh.cpp:4-449 SBVH; seconds per
// scene) and hands CLContext::uploadSceneData two arrays: `Node[]` (48 B, depth-first order, left child = self + 1,
// rightChild / iStart union, parent, nPrims; src/bvhnode.hpp:50-59, flattened by src/sbvh.cpp:52-73) and the u32 index list.
// This builder produces THE SAME FORMAT in a few milliseconds, entirely on the device, so it can stand in for `new SBVH(...)`
// wherever build time matters more than tree quality (interactive edits, first frame):
//
//   1. k_bvh_prims      per-triangle box + centroid, scene centroid bounds (block reduction + ordered-int atomics)
//   2. k_bvh_morton     64-bit keys: 30-bit Morton code of the box centre << 32 | triangle index  (unique => no tie handling)
//   3. radix sort       cub::DeviceRadixSort::SortKeys, bits 0..62 (library call: a plain sort, not a path kernel)
//   4. k_bvh_hierarchy  binary radix tree over the sorted keys, one thread per internal node (Karras 2012)
//   5. k_bvh_fit        bottom-up, second arriver proceeds: boxes, subtree SAH cost with the reference's constants
//                       (costBox = costTri = 1, src/bvh.hpp:72-73; leaf = area * n, inner = 2 * area + children,
//                       src/sbvh.cpp:115,129), collapse to a leaf when n <= maxLeaf and the leaf is not dearer
//                       (the reference's rule, src/sbvh.cpp:133), and the size of every surviving subtree
//   6. k_bvh_emit       every surviving node computes its depth-first index by walking to the root
//                       (left child: +1, right child: +1 + size(left sibling)) and writes its 48-byte record;
//                       leaves of a radix tree cover contiguous key ranges, so the index list IS the sorted key list.
//
// Everything is deterministic (unique keys, fixed association order, no float atomics), so the CPU restatement in
// oracle/bvh_oracle.c reproduces the node and index arrays bit for bit.
#pragma once

#include <cub/device/device_radix_sort.cuh>

#include "flx_device.cuh"

#define FLX_BVH_BLOCK 256

struct BvhBuild
{
    const flx_Triangle *tris;
    uint32_t n;
    uint32_t maxLeaf;
    float triCost;         // SAH cost of a triangle test relative to a box test in the collapse decision (the reference: 1, src/bvh.hpp:72-73)
    unsigned long long *keys, *keysSorted;
    float4 *bmin, *bmax;   // node boxes; node ids: internal i -> i (0 .. n-2), leaf j (sorted position) -> n-1+j
    float4 *primMin, *primMax; // per triangle, by triangle index
    int *parent;           // per node id
    int2 *children;        // per internal node: node ids
    uint2 *range;          // per internal node: first, last sorted position
    float *cost;           // per node id: SAH cost of the subtree as finally built
    uint32_t *size;        // per node id: nodes of the surviving subtree (1 for leaves and collapsed nodes)
    uint32_t *collapsed;   // per internal node
    uint32_t *visits;      // per internal node: arrival counter of the bottom-up pass
    uint32_t *sceneBounds; // 6 ordered ints: centroid min xyz, max xyz
    flx_Node *nodesOut;
    uint32_t *indicesOut;
};

FLX_DEV uint32_t float_to_ordered(float f)
{
    const uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
FLX_DEV float ordered_to_float(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o); }
FLX_DEV float half_area(float4 lo, float4 hi)
{
    const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z;
    return (dx * dy + dy * dz) + dz * dx;
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_bvh_prims(const BvhBuild b)
{
    __shared__ float s_red[6][FLX_BVH_BLOCK / 32];
    const uint32_t i = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    float c[6] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    if (i < b.n)
    {
        const float4 *q = reinterpret_cast<const float4 *>(b.tris + i);
        const float4 p0 = __ldg(q), p1 = __ldg(q + 3), p2 = __ldg(q + 6);
        const float4 lo = make_float4(fminf(fminf(p0.x, p1.x), p2.x), fminf(fminf(p0.y, p1.y), p2.y), fminf(fminf(p0.z, p1.z), p2.z), 0.0f);
        const float4 hi = make_float4(fmaxf(fmaxf(p0.x, p1.x), p2.x), fmaxf(fmaxf(p0.y, p1.y), p2.y), fmaxf(fmaxf(p0.z, p1.z), p2.z), 0.0f);
        b.primMin[i] = lo;
        b.primMax[i] = hi;
        c[0] = c[3] = (lo.x + hi.x) * 0.5f;
        c[1] = c[4] = (lo.y + hi.y) * 0.5f;
        c[2] = c[5] = (lo.z + hi.z) * 0.5f;
    }
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
        float v = c[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
        {
            const float w = __shfl_xor_sync(0xffffffffu, v, o);
            v = k < 3 ? fminf(v, w) : fmaxf(v, w);
        }
        if ((threadIdx.x & 31) == 0)
            s_red[k][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 6)
    {
        const int k = threadIdx.x;
        float v = s_red[k][0];
        for (int w = 1; w < FLX_BVH_BLOCK / 32; w++)
            v = k < 3 ? fminf(v, s_red[k][w]) : fmaxf(v, s_red[k][w]);
        if (k < 3)
            atomicMin(b.sceneBounds + k, float_to_ordered(v));
        else
            atomicMax(b.sceneBounds + k, float_to_ordered(v));
    }
}

FLX_DEV uint32_t spread10(uint32_t v) // 10 bits -> every third bit
{
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
FLX_DEV uint32_t quantize10(float c, float lo, float hi)
{
    const float ext = hi - lo;
    if (!(ext > 0.0f))
        return 0u;
    const float q = ((c - lo) / ext) * 1024.0f;
    return (uint32_t)fminf(fmaxf(q, 0.0f), 1023.0f);
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_bvh_morton(const BvhBuild b)
{
    const uint32_t i = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (i >= b.n)
        return;
    const float lox = ordered_to_float(b.sceneBounds[0]), loy = ordered_to_float(b.sceneBounds[1]), loz = ordered_to_float(b.sceneBounds[2]);
    const float hix = ordered_to_float(b.sceneBounds[3]), hiy = ordered_to_float(b.sceneBounds[4]), hiz = ordered_to_float(b.sceneBounds[5]);
    const float4 lo = b.primMin[i], hi = b.primMax[i];
    const uint32_t x = quantize10((lo.x + hi.x) * 0.5f, lox, hix), y = quantize10((lo.y + hi.y) * 0.5f, loy, hiy), z = quantize10((lo.z + hi.z) * 0.5f, loz, hiz);
    const uint32_t m = (spread10(x) << 2) | (spread10(y) << 1) | spread10(z);
    b.keys[i] = ((unsigned long long)m << 32) | (unsigned long long)i;
}

// length of the common prefix of keys i and j, -1 when j is outside the array (Karras 2012, section 4; keys are unique)
FLX_DEV int key_delta(const unsigned long long *keys, int n, int i, int j)
{
    if (j < 0 || j >= n)
        return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_bvh_hierarchy(const BvhBuild b)
{
    const int n = (int)b.n;
    const int i = (int)(blockIdx.x * FLX_BVH_BLOCK + threadIdx.x);
    if (i >= n - 1)
        return;
    const unsigned long long *k = b.keysSorted;
    const int d = key_delta(k, n, i, i + 1) - key_delta(k, n, i, i - 1) > 0 ? 1 : -1;
    const int dmin = key_delta(k, n, i, i - d);
    int lmax = 2;
    while (key_delta(k, n, i, i + lmax * d) > dmin)
        lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (key_delta(k, n, i, i + (l + t) * d) > dmin)
            l += t;
    const int j = i + l * d;
    const int dnode = key_delta(k, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) // ceil(l / 2), ceil(l / 4), ... , 1
    {
        if (key_delta(k, n, i, i + (s + t) * d) > dnode)
            s += t;
        if (t == 1)
            break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? (n - 1 + gamma) : gamma;
    const int right = (last == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    b.children[i] = make_int2(left, right);
    b.range[i] = make_uint2((uint32_t)first, (uint32_t)last);
    b.parent[left] = i;
    b.parent[right] = i;
    if (i == 0)
        b.parent[0] = -1;
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_bvh_fit(const BvhBuild b)
{
    const uint32_t j = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (j >= b.n)
        return;
    const uint32_t tri = (uint32_t)(b.keysSorted[j] & 0xffffffffull);
    b.indicesOut[j] = tri;
    int node = (int)(b.n - 1 + j);
    {
        const float4 lo = b.primMin[tri], hi = b.primMax[tri];
        b.bmin[node] = lo;
        b.bmax[node] = hi;
        b.cost[node] = half_area(lo, hi) * b.triCost;
        b.size[node] = 1u;
    }
    if (b.n == 1)
        return;
    while (true)
    {
        __threadfence();
        const int p = b.parent[node];
        if (p < 0)
            break;
        if (atomicAdd(b.visits + p, 1u) == 0u)
            break; // the sibling's thread will do the parent
        __threadfence();
        const int2 ch = b.children[p];
        const volatile float4 *vmin = b.bmin, *vmax = b.bmax;
        const float4 lmin = make_float4(vmin[ch.x].x, vmin[ch.x].y, vmin[ch.x].z, 0.0f), lmax = make_float4(vmax[ch.x].x, vmax[ch.x].y, vmax[ch.x].z, 0.0f);
        const float4 rmin = make_float4(vmin[ch.y].x, vmin[ch.y].y, vmin[ch.y].z, 0.0f), rmax = make_float4(vmax[ch.y].x, vmax[ch.y].y, vmax[ch.y].z, 0.0f);
        const float4 lo = make_float4(fminf(lmin.x, rmin.x), fminf(lmin.y, rmin.y), fminf(lmin.z, rmin.z), 0.0f);
        const float4 hi = make_float4(fmaxf(lmax.x, rmax.x), fmaxf(lmax.y, rmax.y), fmaxf(lmax.z, rmax.z), 0.0f);
        const float area = half_area(lo, hi);
        const uint2 r = b.range[p];
        const uint32_t count = r.y - r.x + 1u;
        const volatile float *vcost = b.cost;
        const volatile uint32_t *vsize = b.size;
        const float leafCost = (area * (float)count) * b.triCost;          // parentArea * refs * costTri, sbvh.cpp:129
        const float innerCost = (area * 2.0f + vcost[ch.x]) + vcost[ch.y]; // nodeSAH + children, sbvh.cpp:115,204
        const bool collapse = count <= b.maxLeaf && leafCost <= innerCost;
        b.bmin[p] = lo;
        b.bmax[p] = hi;
        b.cost[p] = collapse ? leafCost : innerCost;
        b.size[p] = collapse ? 1u : 1u + vsize[ch.x] + vsize[ch.y];
        b.collapsed[p] = collapse ? 1u : 0u;
        node = p;
    }
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_bvh_emit(const BvhBuild b)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    const uint32_t total = 2u * b.n - 1u;
    if (id >= total)
        return;
    const bool isLeafNode = id >= b.n - 1u;
    // depth-first index: climb to the root; a collapsed ancestor means this node does not exist in the output
    uint32_t dfs = 0, parentDfsDelta = 0;
    bool first = true;
    int node = (int)id;
    while (true)
    {
        const int p = b.parent[node];
        if (p < 0)
            break;
        if (b.collapsed[p])
            return;
        const int2 ch = b.children[p];
        const uint32_t step = (ch.x == node) ? 1u : 1u + b.size[ch.x];
        if (first)
        {
            parentDfsDelta = step;
            first = false;
        }
        dfs += step;
        node = p;
    }
    flx_Node out;
    const float4 lo = b.bmin[id], hi = b.bmax[id];
    out.bmin.x = lo.x; out.bmin.y = lo.y; out.bmin.z = lo.z; out.bmin.w = 0.0f;
    out.bmax.x = hi.x; out.bmax.y = hi.y; out.bmax.z = hi.z; out.bmax.w = 0.0f;
    out.parent = first ? -1 : (int)(dfs - parentDfsDelta);
    for (int k = 0; k < 7; k++)
        out._pad[k] = 0;
    if (isLeafNode)
    {
        out.iStartOrRightChild = id - (b.n - 1u);
        out.nPrims = 1;
    }
    else if (b.collapsed[id])
    {
        const uint2 r = b.range[id];
        out.iStartOrRightChild = r.x;
        out.nPrims = (uint8_t)(r.y - r.x + 1u);
    }
    else
    {
        out.iStartOrRightChild = dfs + 1u + b.size[b.children[id].x];
        out.nPrims = 0;
    }
    b.nodesOut[dfs] = out;
}

// ================================================================================================ PLOC
// Second builder, for better trees at a few times the build cost: parallel locally-ordered clustering (Meister & Bittner 2018).
// The Morton-sorted triangles start as one cluster each; every round each cluster looks R positions to either side for the
// neighbour whose merged box has the smallest area (ties: the lower position), mutual pairs merge into a new node that takes
// the lower partner's place, the array is compacted (one exclusive sum over (keep, merge) packed in 64 bits), until one cluster
// is left.  SAH cost, collapse decision, surviving-subtree size and triangle count of a node are final the moment it is
// created, so no separate bottom-up pass is needed.  The emit step is the general one (any binary tree): a node's depth-first
// index and the position of its first triangle in the index list come from walking to the root (left child: +1 / +0,
// right child: +1 + size(left sibling) / + triangles(left sibling)).
// Deterministic like the LBVH path: with "lowest position wins ties" the lowest-positioned cluster of a closest pair always has
// a mutual partner, so every round merges at least one pair; oracle/bvh_oracle.c restates it sequentially, bit for bit.
//
// Node ids: leaf j (sorted position) -> j; inner nodes -> n, n+1, ... in order of creation (by round, then by position).
#define FLX_PLOC_RADIUS 16

struct PlocBuild
{
    uint32_t n, maxLeaf;
    float triCost;
    const unsigned long long *keysSorted;
    const float4 *primMin, *primMax;
    float4 *bmin, *bmax;     // per node id (2n - 1)
    int *left, *right, *parent;
    float *cost;
    uint32_t *size, *prims, *collapsed;
    uint32_t *cidA, *cidB;   // cluster arrays (ping-pong): node id at each position
    int *nn;                 // per position: chosen neighbour position
    unsigned long long *flags, *scan; // per position: keep | merge << 32, and its exclusive sum
    uint32_t *depthMax;
    uint32_t *state;         // {clusters left, next node id}: lives on the device so rounds can be enqueued without a host round trip
    flx_Node *nodesOut;
    uint32_t *indicesOut;
};

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ploc_init(const PlocBuild b)
{
    const uint32_t j = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (j >= b.n)
        return;
    const uint32_t tri = (uint32_t)(b.keysSorted[j] & 0xffffffffull);
    const float4 lo = b.primMin[tri], hi = b.primMax[tri];
    b.bmin[j] = lo;
    b.bmax[j] = hi;
    b.left[j] = -1;
    b.right[j] = -1;
    b.parent[j] = -1;
    b.cost[j] = half_area(lo, hi) * b.triCost;
    b.size[j] = 1u;
    b.prims[j] = 1u;
    b.collapsed[j] = 0u;
    b.cidA[j] = j;
    if (j == 0)
    {
        b.state[0] = b.n;
        b.state[1] = b.n;
    }
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ploc_nearest(const PlocBuild b, const uint32_t *cid)
{
    const uint32_t m = b.state[0];
    if (blockIdx.x * FLX_BVH_BLOCK >= m)
        return;
    __shared__ float4 s_lo[FLX_BVH_BLOCK + 2 * FLX_PLOC_RADIUS], s_hi[FLX_BVH_BLOCK + 2 * FLX_PLOC_RADIUS];
    const int base = (int)(blockIdx.x * FLX_BVH_BLOCK) - FLX_PLOC_RADIUS;
    for (int k = threadIdx.x; k < FLX_BVH_BLOCK + 2 * FLX_PLOC_RADIUS; k += FLX_BVH_BLOCK)
    {
        const int q = base + k;
        if (q >= 0 && q < (int)m)
        {
            const uint32_t id = cid[q];
            s_lo[k] = b.bmin[id];
            s_hi[k] = b.bmax[id];
        }
    }
    __syncthreads();
    const int p = (int)(blockIdx.x * FLX_BVH_BLOCK + threadIdx.x);
    if (p >= (int)m)
        return;
    const float4 lo = s_lo[threadIdx.x + FLX_PLOC_RADIUS], hi = s_hi[threadIdx.x + FLX_PLOC_RADIUS];
    float best = 3.402823466e+38f;
    int bestq = -1;
    for (int d = -FLX_PLOC_RADIUS; d <= FLX_PLOC_RADIUS; d++) // ascending position: "<" keeps the lowest position on ties
    {
        const int q = p + d;
        if (d == 0 || q < 0 || q >= (int)m)
            continue;
        const float4 qlo = s_lo[threadIdx.x + FLX_PLOC_RADIUS + d], qhi = s_hi[threadIdx.x + FLX_PLOC_RADIUS + d];
        const float4 ulo = make_float4(fminf(lo.x, qlo.x), fminf(lo.y, qlo.y), fminf(lo.z, qlo.z), 0.0f);
        const float4 uhi = make_float4(fmaxf(hi.x, qhi.x), fmaxf(hi.y, qhi.y), fmaxf(hi.z, qhi.z), 0.0f);
        const float a = half_area(ulo, uhi);
        if (a < best || bestq < 0)
        {
            best = a;
            bestq = q;
        }
    }
    b.nn[p] = bestq;
}

// bound: the number of positions the following scan covers (the cluster count as last known to the host); positions past the
// current count contribute zeros
__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ploc_flags(const PlocBuild b, const uint32_t bound)
{
    const uint32_t m = b.state[0];
    const int p = (int)(blockIdx.x * FLX_BVH_BLOCK + threadIdx.x);
    if (p >= (int)bound)
        return;
    if (p >= (int)m)
    {
        b.flags[p] = 0ull;
        return;
    }
    const int q = b.nn[p];
    const bool mutual = q >= 0 && b.nn[q] == p;
    const unsigned long long keep = (mutual && q < p) ? 0ull : 1ull; // the upper partner of a pair disappears
    const unsigned long long merge = (mutual && p < q) ? 1ull : 0ull; // the lower partner's place takes the new node
    b.flags[p] = keep | (merge << 32);
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ploc_apply(const PlocBuild b, const uint32_t *cid, uint32_t *cidNext)
{
    const uint32_t m = b.state[0], nextId = b.state[1];
    const int p = (int)(blockIdx.x * FLX_BVH_BLOCK + threadIdx.x);
    if (p >= (int)m)
        return;
    const unsigned long long f = b.flags[p], s = b.scan[p];
    if ((f & 1ull) == 0ull)
        return;
    const uint32_t pos = (uint32_t)(s & 0xffffffffull);
    if ((f >> 32) == 0ull)
    {
        cidNext[pos] = cid[p];
        return;
    }
    const uint32_t id = nextId + (uint32_t)(s >> 32);
    const uint32_t l = cid[p], r = cid[b.nn[p]];
    const float4 llo = b.bmin[l], lhi = b.bmax[l], rlo = b.bmin[r], rhi = b.bmax[r];
    const float4 lo = make_float4(fminf(llo.x, rlo.x), fminf(llo.y, rlo.y), fminf(llo.z, rlo.z), 0.0f);
    const float4 hi = make_float4(fmaxf(lhi.x, rhi.x), fmaxf(lhi.y, rhi.y), fmaxf(lhi.z, rhi.z), 0.0f);
    const float area = half_area(lo, hi);
    const uint32_t count = b.prims[l] + b.prims[r];
    const float leafCost = (area * (float)count) * b.triCost;
    const float innerCost = (area * 2.0f + b.cost[l]) + b.cost[r];
    const bool collapse = count <= b.maxLeaf && leafCost <= innerCost;
    b.bmin[id] = lo;
    b.bmax[id] = hi;
    b.left[id] = (int)l;
    b.right[id] = (int)r;
    b.parent[id] = -1;
    b.parent[l] = (int)id;
    b.parent[r] = (int)id;
    b.cost[id] = collapse ? leafCost : innerCost;
    b.size[id] = collapse ? 1u : 1u + b.size[l] + b.size[r];
    b.prims[id] = count;
    b.collapsed[id] = collapse ? 1u : 0u;
    cidNext[pos] = id;
}

// one thread, after k_ploc_apply: the totals of the round's exclusive sum become the new cluster count / next node id
__global__ void k_ploc_advance(const PlocBuild b)
{
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    const uint32_t m = b.state[0];
    const unsigned long long totals = b.scan[m - 1] + b.flags[m - 1];
    b.state[0] = (uint32_t)(totals & 0xffffffffull);
    b.state[1] += (uint32_t)(totals >> 32);
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ploc_emit(const PlocBuild b)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    const uint32_t total = 2u * b.n - 1u;
    if (id >= total)
        return;
    uint32_t dfs = 0, firstPrim = 0, parentDelta = 0, depth = 0;
    bool atRoot = true, alive = true;
    int node = (int)id;
    while (true)
    {
        const int p = b.parent[node];
        if (p < 0)
            break;
        if (b.collapsed[p])
            alive = false; // not a node of the output; its triangle still needs its place in the index list
        const int l = b.left[p];
        const bool isLeft = l == node;
        const uint32_t step = isLeft ? 1u : 1u + b.size[l];
        if (atRoot)
        {
            parentDelta = step;
            atRoot = false;
        }
        dfs += step;
        firstPrim += isLeft ? 0u : b.prims[l];
        depth++;
        node = p;
    }
    if (id < b.n) // a triangle: its place in the index list is its in-order rank, whether or not its leaf node survives
        b.indicesOut[firstPrim] = (uint32_t)(b.keysSorted[id] & 0xffffffffull);
    if (!alive)
        return;
    atomicMax(b.depthMax, depth);
    flx_Node out;
    const float4 lo = b.bmin[id], hi = b.bmax[id];
    out.bmin.x = lo.x; out.bmin.y = lo.y; out.bmin.z = lo.z; out.bmin.w = 0.0f;
    out.bmax.x = hi.x; out.bmax.y = hi.y; out.bmax.z = hi.z; out.bmax.w = 0.0f;
    out.parent = atRoot ? -1 : (int)(dfs - parentDelta);
    for (int k = 0; k < 7; k++)
        out._pad[k] = 0;
    if (id < b.n || b.collapsed[id])
    {
        out.iStartOrRightChild = firstPrim;
        out.nPrims = (uint8_t)b.prims[id];
    }
    else
    {
        out.iStartOrRightChild = dfs + 1u + b.size[b.left[id]];
        out.nPrims = 0;
    }
    b.nodesOut[dfs] = out;
}

// ================================================================================================ parallel reinsertion
// FLX_BVH_PLOC_OPT: a post-pass over the finished PLOC tree, after Meister & Bittner, "Parallel reinsertion for bounding volume
// hierarchy optimization" (2018).  What the reference's SBVH gets from spatial splits -- little overlap between siblings
// (src/sbvh.cpp:118-142) -- a bottom-up builder can approach by moving subtrees to where they enlarge the fewest boxes: on Conference
// the PLOC tree makes a ray test both children of a node 50 % more often than the SBVH does, which is where its 10 % traversal deficit
// came from.  Per iteration, on the tree as it stands:
//   k_ri_search  every node x (not the root, not a child of the root, not inside a collapsed leaf) looks for the node y next to which it
//                would sit best: x and its parent p leave (the sibling s takes p's place), p comes back as the parent of (y, x).  Gain =
//                the half-area inner nodes lose (p itself; the ancestors of p below the lowest common ancestor "pivot" shrink) minus what
//                they gain (p's new box; the ancestors of y below the pivot grow).  The pivot walks from p to the root; under each pivot
//                the subtree on the other side is searched depth first, left child first, pruned where even a perfect fit (direct cost =
//                the area of x) cannot beat the best gain so far.  Read-only, one thread per node;
//   k_ri_lock    a move with gain > 0 puts (gain bits << 32 | x) on the six nodes whose links it rewrites -- x, p, s, the grandparent, y
//                and y's parent -- with atomicMax: the largest gain wins a contested node, whatever the order of arrival;
//   k_ri_cand    a move that holds all six is a candidate;
//   k_ri_guard   simultaneous moves with disjoint link sets could still close a cycle (x1 goes below x2 while x2 goes below x1, or a longer
//                chain): that takes, for every move of the chain, another moving node strictly between its pivot and its target -- so a
//                candidate that finds another candidate's x on its way from y up to the pivot stands back;
//   k_ri_apply   the remaining moves rewrite their links (disjoint sets: no two write the same node);
//   k_ri_refit   boxes bottom-up, second arriver proceeds (min / max: exact, order-free).
// After the last iteration the same bottom-up pass also recomputes SAH cost, subtree size, triangle count and the collapse decision
// with the builder's formulas, and k_ploc_emit writes the arrays as before.  No float sums across threads, ties broken by node id:
// oracle/bvh_oracle.c (ploc_reinsert) restates the iterations sequentially and arrives at the same arrays bit for bit.
#define FLX_RI_STACK 128

struct ReinsertView
{
    PlocBuild b;
    int root;
    uint32_t total;
    uint32_t *alive;            // per node id: 0 = inside a collapsed leaf, 1 = inner node of the output tree, 2 = leaf of the output tree
    float *gain;
    int *out, *pivot;           // per node id: best target (-1: stay) and the pivot it was found under
    unsigned long long *lock;
    uint32_t *cand, *win;
    uint32_t *visits;           // arrival counters of the bottom-up pass
    uint32_t *moves;            // moves carried out, all iterations
};

FLX_DEV bool ri_is_leaf(const PlocBuild &b, int id) { return (uint32_t)id < b.n || b.collapsed[id] != 0u; }
FLX_DEV void ri_union(float4 alo, float4 ahi, float4 blo, float4 bhi, float4 &lo, float4 &hi)
{
    lo = make_float4(fminf(alo.x, blo.x), fminf(alo.y, blo.y), fminf(alo.z, blo.z), 0.0f);
    hi = make_float4(fmaxf(ahi.x, bhi.x), fmaxf(ahi.y, bhi.y), fmaxf(ahi.z, bhi.z), 0.0f);
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_alive(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total)
        return;
    uint32_t alive = 1u;
    for (int p = r.b.parent[id]; p >= 0; p = r.b.parent[p])
        if (r.b.collapsed[p])
        {
            alive = 0u;
            break;
        }
    r.alive[id] = alive ? (ri_is_leaf(r.b, (int)id) ? 2u : 1u) : 0u;
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_search(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total)
        return;
    const PlocBuild &b = r.b;
    r.lock[id] = 0ull;
    r.cand[id] = 0u;
    r.win[id] = 0u;
    r.visits[id] = 0u;
    const int x = (int)id;
    float best = 0.0f;
    int bestOut = -1, bestPivot = -1;
    const int p = r.alive[id] ? b.parent[x] : -1;
    if (x != r.root && p >= 0 && p != r.root)
    {
        const float4 xlo = b.bmin[x], xhi = b.bmax[x];
        const float aX = half_area(xlo, xhi);
        float dDec = half_area(b.bmin[p], b.bmax[p]);
        float4 pathLo = xlo, pathHi = xhi; // set when the pivot leaves p
        int child = x, pivot = p;
        int stackNode[FLX_RI_STACK];
        float stackInc[FLX_RI_STACK];
        while (true)
        {
            const int l = b.left[pivot];
            const int other = l == child ? b.right[pivot] : l;
            int sp = 0;
            stackNode[sp] = other;
            stackInc[sp] = 0.0f;
            sp++;
            while (sp > 0)
            {
                --sp;
                const int y = stackNode[sp];
                const float inc = stackInc[sp];
                const float4 ylo = b.bmin[y], yhi = b.bmax[y];
                float4 ulo, uhi;
                ri_union(ylo, yhi, xlo, xhi, ulo, uhi);
                const float direct = half_area(ulo, uhi);
                const float gain = (dDec - inc) - direct;
                if (gain > best)
                {
                    best = gain;
                    bestOut = y;
                    bestPivot = pivot;
                }
                if (!ri_is_leaf(b, y))
                {
                    const float incChild = (inc + direct) - half_area(ylo, yhi);
                    if (((dDec - incChild) - aX) > best && sp + 2 <= FLX_RI_STACK)
                    {
                        stackNode[sp] = b.right[y];
                        stackInc[sp] = incChild;
                        sp++;
                        stackNode[sp] = b.left[y];
                        stackInc[sp] = incChild;
                        sp++;
                    }
                }
            }
            if (pivot == r.root)
                break;
            if (pivot == p)
            {
                pathLo = b.bmin[other];
                pathHi = b.bmax[other];
            }
            else
            {
                ri_union(pathLo, pathHi, b.bmin[other], b.bmax[other], pathLo, pathHi);
                dDec = dDec + (half_area(b.bmin[pivot], b.bmax[pivot]) - half_area(pathLo, pathHi));
            }
            child = pivot;
            pivot = b.parent[pivot];
        }
    }
    r.gain[id] = best;
    r.out[id] = bestOut;
    r.pivot[id] = bestPivot;
}

// the six nodes whose links the move of x next to y rewrites
FLX_DEV void ri_link_set(const PlocBuild &b, int x, int y, int (&v)[6])
{
    const int p = b.parent[x];
    v[0] = x;
    v[1] = p;
    v[2] = b.left[p] == x ? b.right[p] : b.left[p];
    v[3] = b.parent[p];
    v[4] = y;
    v[5] = b.parent[y];
}
FLX_DEV unsigned long long ri_key(float gain, uint32_t x) { return ((unsigned long long)__float_as_uint(gain) << 32) | (unsigned long long)x; }

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_lock(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total || r.out[id] < 0)
        return;
    int v[6];
    ri_link_set(r.b, (int)id, r.out[id], v);
    const unsigned long long key = ri_key(r.gain[id], id);
#pragma unroll
    for (int k = 0; k < 6; k++)
        atomicMax(r.lock + v[k], key);
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_cand(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total || r.out[id] < 0)
        return;
    int v[6];
    ri_link_set(r.b, (int)id, r.out[id], v);
    const unsigned long long key = ri_key(r.gain[id], id);
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 6; k++)
        ok = ok && r.lock[v[k]] == key;
    r.cand[id] = ok ? 1u : 0u;
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_guard(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total || !r.cand[id])
        return;
    bool ok = true;
    const int pivot = r.pivot[id];
    for (int v = r.b.parent[r.out[id]]; v != pivot; v = r.b.parent[v])
        if (r.cand[v])
            ok = false;
    r.win[id] = ok ? 1u : 0u;
}

__global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_apply(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total || !r.win[id])
        return;
    const PlocBuild &b = r.b;
    const int x = (int)id, y = r.out[id];
    const int p = b.parent[x], s = b.left[p] == x ? b.right[p] : b.left[p], g = b.parent[p];
    if (b.left[g] == p)
        b.left[g] = s;
    else
        b.right[g] = s;
    b.parent[s] = g;
    const int yp = b.parent[y]; // may be g: read after g's link to p has become the link to s
    if (b.left[yp] == y)
        b.left[yp] = p;
    else
        b.right[yp] = p;
    b.parent[p] = yp;
    b.left[p] = y;
    b.right[p] = x;
    b.parent[y] = p;
    atomicAdd(r.moves, 1u);
}

// bottom-up over the output tree from its leaves (single triangles and collapsed nodes); FULL: also SAH cost, surviving-subtree size,
// triangle count and the collapse decision, as k_ploc_apply computes them when it creates a node
template <bool FULL> __global__ void __launch_bounds__(FLX_BVH_BLOCK) k_ri_refit(const ReinsertView r)
{
    const uint32_t id = blockIdx.x * FLX_BVH_BLOCK + threadIdx.x;
    if (id >= r.total)
        return;
    const PlocBuild &b = r.b;
    if (r.alive[id] != 2u) // (not b.collapsed: the FULL pass rewrites it while other threads are still starting)
        return;
    int node = (int)id;
    while (true)
    {
        __threadfence();
        const int p = b.parent[node];
        if (p < 0)
            break;
        if (atomicAdd(r.visits + p, 1u) == 0u)
            break; // the sibling's thread will do the parent
        __threadfence();
        const int l = b.left[p], rr = b.right[p];
        const volatile float4 *vmin = b.bmin, *vmax = b.bmax;
        const float4 lmin = make_float4(vmin[l].x, vmin[l].y, vmin[l].z, 0.0f), lmax = make_float4(vmax[l].x, vmax[l].y, vmax[l].z, 0.0f);
        const float4 rmin = make_float4(vmin[rr].x, vmin[rr].y, vmin[rr].z, 0.0f), rmax = make_float4(vmax[rr].x, vmax[rr].y, vmax[rr].z, 0.0f);
        float4 lo, hi;
        ri_union(lmin, lmax, rmin, rmax, lo, hi);
        b.bmin[p] = lo;
        b.bmax[p] = hi;
        if (FULL)
        {
            const volatile float *vcost = b.cost;
            const volatile uint32_t *vsize = b.size, *vprims = b.prims;
            const float area = half_area(lo, hi);
            const uint32_t count = vprims[l] + vprims[rr];
            const float leafCost = (area * (float)count) * b.triCost;
            const float innerCost = (area * 2.0f + vcost[l]) + vcost[rr];
            const bool collapse = count <= b.maxLeaf && leafCost <= innerCost;
            b.cost[p] = collapse ? leafCost : innerCost;
            b.size[p] = collapse ? 1u : 1u + vsize[l] + vsize[rr];
            b.prims[p] = count;
            b.collapsed[p] = collapse ? 1u : 0u;
        }
        node = p;
    }
}
