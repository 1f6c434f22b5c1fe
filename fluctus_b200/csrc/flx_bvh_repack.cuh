// flx_bvh_repack.cuh -- reference hierarchy (Node[] 48 B in depth-first order + index list + Triangle[] 160 B; src/bvhnode.hpp:50-59,
// src/sbvh.cpp:52-73) -> the traversal layout of flx_trace.cuh (TNode 64 B per inner node, TTri 64 B per leaf reference), on the
// device.  Same output, bit for bit, as the host version `repackBvh` in flx_api.cu (kept as the checker, FLX_TUNE_REPACK_ON_HOST);
// the host only picks the top-of-tree treelet (a 4096-step priority queue).  SURVEY 8(f-1), "GPU repacker".
//
//   k_repack_flags   per node: validate links / index ranges, flag = inner node outside the treelet, prims = leaf size
//   exclusive sums   (cub::DeviceScan) over both -> position of every unplaced inner node / first TTri of every leaf
//   k_repack_emit    per node: inner -> its TNode (both child boxes + child references), leaf -> its TTri run
#pragma once

#include <cub/device/device_scan.cuh>

#include "flx_device.cuh"

struct RepackView
{
    const flx_Node *nodes;
    const uint32_t *indices;
    const flx_Triangle *tris;
    uint32_t nNodes, nIndices, nTris, treeletCount;
    const int *treeletPos;    // per node: position in the treelet, -1 = not in it
    uint32_t *flagInner;      // per node: 1 = inner node outside the treelet
    uint32_t *leafPrims;      // per node: nPrims (0 for inner nodes)
    const uint32_t *scanInner, *scanLeaf; // exclusive sums of the two arrays above
    float4 *tnodes, *ttris;
    uint32_t *error;          // 0 = fine; else (kind << 28 | node index + 1), first writer wins
};

FLX_DEV void repack_fail(uint32_t *error, uint32_t kind, uint32_t node) { atomicCAS(error, 0u, (kind << 28) | ((node + 1u) & 0x0fffffffu)); }

__global__ void __launch_bounds__(256) k_repack_scatter_treelet(const uint32_t *treelet, uint32_t count, int *treeletPos)
{
    const uint32_t k = blockIdx.x * 256 + threadIdx.x;
    if (k < count)
        treeletPos[treelet[k]] = (int)k;
}

__global__ void __launch_bounds__(256) k_repack_flags(const RepackView r)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= r.nNodes)
        return;
    const uint32_t link = r.nodes[i].iStartOrRightChild, n = r.nodes[i].nPrims;
    if (n == 0)
    {
        if (i + 1 >= r.nNodes || link >= r.nNodes || link <= i + 1)
            repack_fail(r.error, 1u, i);
        r.flagInner[i] = r.treeletPos[i] < 0 ? 1u : 0u;
        r.leafPrims[i] = 0u;
    }
    else
    {
        if ((unsigned long long)link + n > r.nIndices)
            repack_fail(r.error, 2u, i);
        r.flagInner[i] = 0u;
        r.leafPrims[i] = n;
    }
}

FLX_DEV int repack_ref(const RepackView &r, uint32_t i) // child reference of reference node i
{
    if (r.nodes[i].nPrims != 0)
        return ~(int)r.scanLeaf[i];
    const int t = r.treeletPos[i];
    return t >= 0 ? t : (int)(r.treeletCount + r.scanInner[i]);
}

__global__ void __launch_bounds__(256) k_repack_emit(const RepackView r)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= r.nNodes || *r.error)
        return;
    const flx_Node &nd = r.nodes[i];
    if (nd.nPrims == 0)
    {
        const uint32_t li = i + 1, ri = nd.iStartOrRightChild;
        const flx_Node &L = r.nodes[li], &R = r.nodes[ri];
        float4 *q = r.tnodes + (size_t)repack_ref(r, i) * 4;
        q[0] = make_float4(L.bmin.x, L.bmin.y, L.bmin.z, L.bmax.x);
        q[1] = make_float4(L.bmax.y, L.bmax.z, R.bmin.x, R.bmin.y);
        q[2] = make_float4(R.bmin.z, R.bmax.x, R.bmax.y, R.bmax.z);
        q[3] = make_float4(__int_as_float(repack_ref(r, li)), __int_as_float(repack_ref(r, ri)), 0.0f, 0.0f);
    }
    else
    {
        const uint32_t s = nd.iStartOrRightChild, n = nd.nPrims;
        float4 *q = r.ttris + (size_t)r.scanLeaf[i] * 4;
        for (uint32_t k = 0; k < n; k++, q += 4)
        {
            const uint32_t ti = r.indices[s + k];
            if (ti >= r.nTris || ti > 0x7fffffffu)
            {
                repack_fail(r.error, 3u, i);
                return;
            }
            const flx_Triangle &T = r.tris[ti];
            const uint32_t tag = ti | (k + 1 == n ? 0x80000000u : 0u);
            // the float differences below are the ones the reference forms per test (intersect.cl:66-67)
            q[0] = make_float4(T.v0.p.x, T.v0.p.y, T.v0.p.z, __uint_as_float(tag));
            q[1] = make_float4(T.v1.p.x - T.v0.p.x, T.v1.p.y - T.v0.p.y, T.v1.p.z - T.v0.p.z, 0.0f);
            q[2] = make_float4(T.v2.p.x - T.v0.p.x, T.v2.p.y - T.v0.p.y, T.v2.p.z - T.v0.p.z, 0.0f);
            q[3] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        }
    }
}

// TAttr (flx_trace.cuh): the shading attributes of every triangle, gathered from the 160-byte records into 64 bytes.  Same values, so the
// hit record is the same bit for bit; the extension kernel's write-back reads two L1 wavefronts per hit instead of seven.
__global__ void __launch_bounds__(256) k_build_tattr(const flx_Triangle *tris, uint32_t nTris, float4 *attr)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i >= nTris)
        return;
    const flx_Triangle &T = tris[i];
    float4 *q = attr + (size_t)i * 4;
    q[0] = make_float4(T.v0.n.x, T.v0.n.y, T.v0.n.z, T.v1.n.x);
    q[1] = make_float4(T.v1.n.y, T.v1.n.z, T.v2.n.x, T.v2.n.y);
    q[2] = make_float4(T.v2.n.z, T.v0.t.x, T.v0.t.y, T.v1.t.x);
    q[3] = make_float4(T.v1.t.y, T.v2.t.x, T.v2.t.y, __int_as_float(T.matId));
}
