// flx_trace.cuh -- SBVH traversal for the extension (closest-hit) and shadow (any-hit) stages.
//
// Replaces bvh_intersect / bvh_occluded / intersectAABB / intersectTriangle / intersectLight
// of the reference (src/bvh.cl:234-373, src/intersect.cl:41-155) and the two kernels that call
// them (src/wf_extrays.cl:5-36, src/wf_shadowrays.cl:6-37).
//
// Device layout (built once per scene by repack_bvh in flx_api.cu from the reference's 48-byte
// DFS node array, its index list and its 160-byte triangles):
//
//   TNode, 64 B, one per INNER reference node -- both child boxes and both child references in
//   one 64-byte record (half a 128-B line, four LDG.128):
//       q0 = (Lmin.x, Lmin.y, Lmin.z, Lmax.x)   q1 = (Lmax.y, Lmax.z, Rmin.x, Rmin.y)
//       q2 = (Rmin.z, Rmax.x, Rmax.y, Rmax.z)   q3 = (leftRef, rightRef, -, -)  as int
//     a child reference >= 0 is a TNode index; < 0 is ~(first TTri of a leaf).
//     Node ORDER: a treelet of the nodes with the largest box areas, grown greedily from the root, comes first (it is
//     what the TOP traversal variant stages in shared memory), the remaining nodes follow in DFS order.
//   TTri, 64 B (two 256-bit loads), one per leaf REFERENCE (duplicated SBVH references stay duplicated, so a leaf
//   is one contiguous run):
//       q0 = (v0.x, v0.y, v0.z, bits(triangle index | last-in-leaf << 31))
//       q1 = (v1 - v0, 0)   q2 = (v2 - v0, 0)   q3 = unused
//     the edges are the exact float differences the reference computes per test
//     (intersect.cl:66-67), hoisted to build time: same bits, six fewer subtractions per test.
//
// The visiting order, the box test, the Moeller-Trumbore test and every comparison are the
// reference's, so the closest hit (t, u, v, triangle) is bit-identical, ties included.  Shading
// attributes (normals, uv, matId) are fetched from the 160-byte triangle once per ray after
// traversal instead of at every closest-hit update (bvh.cl:273-278): same values, because they
// depend only on the final (triangle, u, v).
#pragma once

#include "flx_device.cuh"

#define FLX_STACK_DEPTH 64 // reference: uint stack[64] (bvh.cl:240); builders cap depth at 64 (bvh.hpp:71)

// Per-ray work counters in the reference's terms: V = nodes popped (inner or leaf, bvh.cl:251/329),
// B = child boxes tested (bvh.cl:283-284), T = triangles tested (bvh.cl:260), U = closest-hit updates (bvh.cl:271-279).
// They feed the algorithmic-bytes numerator of the roofline; NoCount compiles to nothing.
struct NoCount
{
    FLX_DEV void inner() {}
    FLX_DEV void leaf() {}
    FLX_DEV void tri() {}
    FLX_DEV void update() {}
    FLX_DEV void hit() {}
    FLX_DEV void leafEnd() {}
};
struct RayCount
{
    unsigned V = 0, B = 0, T = 0, U = 0, leafImproved = 0;
    FLX_DEV void inner() { V++; B += 2; }
    FLX_DEV void leaf() { V++; }
    FLX_DEV void tri() { T++; }
    FLX_DEV void update() { U++; }
    // flx_trace_greedy.cuh folds a leaf's best into the ray's best triangle by triangle; U stays "leaves that improved the hit"
    FLX_DEV void hit() { leafImproved = 1; }
    FLX_DEV void leafEnd() { U += leafImproved; leafImproved = 0; }
};

struct BvhView
{
    const float4 *nodes; // TNode as 4 x float4
    const float4 *tris;  // TTri as 4 x float4
    const float4 *attr;  // TAttr as 4 x float4, one per TRIANGLE: what a closest hit needs for its hit record -- the three vertex normals, the
                         // three texture coordinates and the material index (n0.xyz n1.x | n1.yz n2.xy | n2.z t0.xy t1.x | t1.y t2.xy matId) --
                         // in 64 bytes = two 256-bit loads instead of seven 128-bit loads scattered over the 160-byte triangle
    int rootRef;
};

struct F8
{
    float v[8];
};
FLX_DEV F8 ldg256(const void *p) // 32-byte aligned, read-only path: LDG.E.256, new on sm_100
{
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                 : "l"(p));
    return r;
}

// The same load with an L1 eviction-priority hint (SASS: LDG.E.EL / .EF / .NA): 1 = evict last, 2 = evict first, 3 = do not allocate in L1.
// The traversal lives on L1: inner nodes are revisited by every ray and are asked to stay (FLX_HINT_NODE = evict last); leaf
// triangles are touched once per visit and would push nodes and stack lines out, so they are the first to go (FLX_HINT_TRI = evict first);
// hit attributes are read once per ray and bypass L1 (FLX_HINT_ATTR = no allocation).  Measured (profiles/r2_cache_hints.txt): with the
// triangles evict-first Conference 2984 -> 3018, Country Kitchen 3018 -> 3025, Luxball 3638 -> 3634 Mrays/s.  Triangles WITHOUT allocation are
// another 0.8 % on Conference but cost 15 % on Luxball (3090 Mrays/s), whose coherent rays re-read a leaf's triangles from L1 -- not taken.
#ifndef FLX_HINT_NODE
#define FLX_HINT_NODE 1
#endif
#ifndef FLX_HINT_TRI
#define FLX_HINT_TRI 2
#endif
#ifndef FLX_HINT_ATTR
#define FLX_HINT_ATTR 3
#endif
template <int HINT> FLX_DEV F8 ldg256_hint(const void *p)
{
    F8 r;
    if (HINT == 1)
        asm volatile("ld.global.nc.L1::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                     : "l"(p));
    else if (HINT == 2)
        asm volatile("ld.global.nc.L1::evict_first.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                     : "l"(p));
    else if (HINT == 3)
        asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
                     : "l"(p));
    else
        r = ldg256(p);
    return r;
}

// The attributes of a closest hit (bvh.cl:273-278: interpolated, normalised vertex normal; interpolated texture coordinates; material
// index) from the 64-byte TAttr record of the winning triangle.  The reference interpolates float3 texture coordinates and keeps .xy; z does
// not reach the result, so it is not stored.
FLX_DEV void hit_attributes(const BvhView &bvh, int tri, float ub, float vb, V3 &N, float &tu, float &tv, int &matId)
{
    const float4 *q = bvh.attr + 4 * (size_t)tri;
    const F8 a = ldg256_hint<FLX_HINT_ATTR>(q), b = ldg256_hint<FLX_HINT_ATTR>(q + 2);
    N = norm3(bary3(ub, vb, v3(a.v[0], a.v[1], a.v[2]), v3(a.v[3], a.v[4], a.v[5]), v3(a.v[6], a.v[7], b.v[0])));
    const V3 uv = bary3(ub, vb, v3(b.v[1], b.v[2], 0.0f), v3(b.v[3], b.v[4], 0.0f), v3(b.v[5], b.v[6], 0.0f));
    tu = uv.x;
    tv = uv.y;
    matId = __float_as_int(b.v[7]);
}

// slab test of one child box (reference: intersectAABB, src/intersect.cl:41-60)
FLX_DEV bool box_test(float bminx, float bminy, float bminz, float bmaxx, float bmaxy, float bmaxz, V3 o, V3 idir, float tprev, float &tnear)
{
    const float ax = (bminx - o.x) * idir.x, ay = (bminy - o.y) * idir.y, az = (bminz - o.z) * idir.z;
    const float bx = (bmaxx - o.x) * idir.x, by = (bmaxy - o.y) * idir.y, bz = (bmaxz - o.z) * idir.z;
    const float tmin = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
    const float tmax = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tnear = tmin;
    if (tmax < 0.0f)
        return false;
    if (tmin > tmax)
        return false;
    return tmin < tprev;
}

// Moeller-Trumbore with precomputed edges (reference: intersectTriangle, src/intersect.cl:63-93)
FLX_DEV bool tri_test(V3 v0, V3 s1, V3 s2, V3 o, V3 d, float &t, float &u, float &v)
{
    const V3 pvec = cross3(d, s2);
    const float det = dot3(s1, pvec);
    if (fabsf(det) < 1e-12f)
        return false;
    const float iDet = 1.0f / det;
    const V3 tvec = o - v0;
    u = dot3(tvec, pvec) * iDet;
    if (u < 0.0f || u > 1.0f)
        return false;
    const V3 qvec = cross3(tvec, s1);
    v = dot3(d, qvec) * iDet;
    if (v < 0.0f || u + v > 1.0f)
        return false;
    t = dot3(s2, qvec) * iDet;
    if (t < 0.0f)
        return false;
    return true;
}

// Closest hit. On return tbest/ubest/vbest/tribest describe the hit (tribest = -1: none).
// STACK is any int-indexable object (local array or strided shared-memory view).
template <class STACK, class COUNT> FLX_DEV void trace_closest(const BvhView &bvh, V3 o, V3 d, float &tbest, float &ubest, float &vbest, int &tribest, STACK &stack, COUNT &cnt)
{
    const V3 idir = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z); // native_recip pinned to IEEE 1/x
    int sp = 0;
    int cur = bvh.rootRef;
    while (true)
    {
        if (cur >= 0)
        {
            cnt.inner();
            const float4 *n = bvh.nodes + 4 * (size_t)cur;
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2);
            const int4 q3 = __ldg(reinterpret_cast<const int4 *>(n + 3));
            float ln, rn;
            const bool lh = box_test(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, idir, tbest, ln);
            const bool rh = box_test(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, idir, tbest, rn);
            if (lh && rh)
            {
                int nearRef = q3.x, farRef = q3.y;
                if (rn < ln) // right child closer -> visit it first (bvh.cl:292); ties keep left first
                {
                    nearRef = q3.y;
                    farRef = q3.x;
                }
                stack[sp++] = farRef;
                cur = nearRef;
                continue;
            }
            if (lh)
            {
                cur = q3.x;
                continue;
            }
            if (rh)
            {
                cur = q3.y;
                continue;
            }
        }
        else
        {
            // leaf: best of the leaf first, then strict "<" against the ray's best (bvh.cl:255-279)
            float tmin = 3.402823466e+38f, umin = 0.0f, vmin = 0.0f;
            int imin = -1;
            cnt.leaf();
            const float4 *p = bvh.tris + 4 * (size_t)(~cur);
            while (true)
            {
                const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                const int tag = __float_as_int(a.w);
                float t, u, v;
                cnt.tri();
                if (tri_test(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), o, d, t, u, v))
                {
                    if (t > 0.0f && t < tmin)
                    {
                        imin = tag & 0x7fffffff;
                        tmin = t;
                        umin = u;
                        vmin = v;
                    }
                }
                if (tag < 0)
                    break;
                p += 4;
            }
            if (imin != -1 && tmin < tbest)
            {
                cnt.update();
                tribest = imin;
                tbest = tmin;
                ubest = umin;
                vbest = vmin;
            }
        }
        if (sp == 0)
            break;
        cur = stack[--sp];
    }
}

// Any hit with 0 < t < maxDist (reference: bvh_occluded, src/bvh.cl:312-373). The answer does not
// depend on the visiting order; near-first is kept because it finds occluders soonest.
template <class STACK, class COUNT> FLX_DEV bool trace_any(const BvhView &bvh, V3 o, V3 d, float maxDist, STACK &stack, COUNT &cnt)
{
    const V3 idir = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    int sp = 0;
    int cur = bvh.rootRef;
    while (true)
    {
        if (cur >= 0)
        {
            cnt.inner();
            const float4 *n = bvh.nodes + 4 * (size_t)cur;
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2);
            const int4 q3 = __ldg(reinterpret_cast<const int4 *>(n + 3));
            float ln, rn;
            const bool lh = box_test(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, idir, maxDist, ln);
            const bool rh = box_test(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, idir, maxDist, rn);
            if (lh && rh)
            {
                int nearRef = q3.x, farRef = q3.y;
                if (rn < ln)
                {
                    nearRef = q3.y;
                    farRef = q3.x;
                }
                stack[sp++] = farRef;
                cur = nearRef;
                continue;
            }
            if (lh)
            {
                cur = q3.x;
                continue;
            }
            if (rh)
            {
                cur = q3.y;
                continue;
            }
        }
        else
        {
            cnt.leaf();
            const float4 *p = bvh.tris + 4 * (size_t)(~cur);
            while (true)
            {
                const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
                float t, u, v;
                cnt.tri();
                if (tri_test(v3(a.x, a.y, a.z), v3(b.x, b.y, b.z), v3(c.x, c.y, c.z), o, d, t, u, v) && t > 0.0f && t < maxDist)
                    return true;
                if (__float_as_int(a.w) < 0)
                    break;
                p += 4;
            }
        }
        if (sp == 0)
            break;
        cur = stack[--sp];
    }
    return false;
}

// The area-light quad as two triangles (reference: intersectTriangleLocal + intersectLight,
// src/intersect.cl:96-155). Returns true and lowers tres when the quad is hit in front of tres.
FLX_DEV bool light_tri(V3 p0, V3 p1, V3 p2, V3 o, V3 d, float &tres)
{
    float t, u, v;
    if (!tri_test(p0, p1 - p0, p2 - p0, o, d, t, u, v))
        return false;
    if (t > tres) // t < 0 already rejected by tri_test
        return false;
    tres = t;
    return true;
}
FLX_DEV bool light_quad(const flx_AreaLight &L, V3 o, V3 d, float &tres)
{
    const V3 N = v3(L.N), pos = v3(L.pos), right = v3(L.right), up = v3(L.up);
    if (dot3(d, N) > 0.0f)
        return false; // back side
    const V3 tl = (pos + L.size.x * right) + L.size.y * up;
    const V3 tr = (pos - L.size.x * right) + L.size.y * up;
    const V3 bl = (pos + L.size.x * right) - L.size.y * up;
    const V3 br = (pos - L.size.x * right) - L.size.y * up;
    const bool first = light_tri(tl, bl, br, o, d, tres);
    const bool second = light_tri(tl, br, tr, o, d, tres);
    return first || second;
}
