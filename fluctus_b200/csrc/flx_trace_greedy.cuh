// flx_trace_greedy.cuh -- tuning variant 3 of the traversal stage: persistent threads, dynamic ray fetch, and a warp that at
// every iteration runs ONE step of the ONE kind of work most of its lanes are waiting for.  An experiment of round 2, kept
// because it is bit-exact and instructive; the production kernel stays flx_trace_persistent.cuh (variant 1).
//
// Motivation.  ncu --import-source on the metric workload (profiles/r2_base_trace_regions.txt) splits the round-1 kernel's
// issued instructions like this:
//
//     region                         share of issued instructions   lanes active (of 32)
//     inner-node loop header + vote            9 %                     32   (pure overhead)
//     inner-node step (load, slabs, descend)  50 %                     17 - 22
//     leaf loop (all triangles of one leaf)   26 %                      6 - 10   (shadow: 6.4 falling to 3.4)
//     fetch / write-back / outer control      15 %
//
// The leaf loop looks like the hole: "while-while" hands every lane that reached a leaf its WHOLE leaf (1..8 triangles, 3.3 on
// average), so the phase lasts as long as the largest leaf among the lanes.  Here a lane is in one of two working states --
// I: at an inner node, T: at triangle k of a leaf -- and one loop iteration is one inner-node step or ONE branch-free triangle
// test for all lanes in that state, chosen by majority (popc of two ballots, `innerBias` shifts the balance).  No per-leaf state
// has to survive the switch because the reference's "best of the leaf, then strict < against the ray's best" (src/bvh.cl:255-279)
// folds into the ray's best directly:
//
//     reference:  tmin = FLT_MAX; for each tri: if (t > 0 && t < tmin) tmin = t ...;  if (imin != -1 && tmin < hit.t) hit = ...
//     here:       for each tri: if (t > 0 && t < tbest) tbest = t ...
//
// Same result in every case: both keep the FIRST triangle with the smallest t below the old best, and neither replaces on a
// tie with the old best.  The leaf cursor is `cur` itself (a leaf reference is ~(TTri index); the next triangle is cur - 1).
//
// Result (profiles/r2_trace_variants.txt): bit-identical path state at every size, and the SAME speed -- 2738 vs 2768 Mrays/s,
// extension kernel alone 0.716 vs 0.712 ms.  Triangle steps do run fuller, but majority rule drives the two populations
// towards 50/50, so inner-node steps (two thirds of the work) run emptier than under while-while, which drains one state
// before it turns to the other.  And the ceiling is lower than the lane count suggests: the kernel sits at 73 % of the issue
// slots AND 73 % of the L1 data pipe (l1tex__data_pipe_lsu_wavefronts: every lane's 32-byte load of a divergent address is its
// own wavefront, two per node or triangle record) -- packing lanes better removes issued instructions but not one wavefront.
// DESIGN.md 4.1 has the numbers.
#pragma once

#include "flx_trace_persistent.cuh"

// Moeller-Trumbore with precomputed edges, without early returns (reference: intersectTriangle, src/intersect.cl:63-93).
// Every intermediate is computed; a degenerate determinant makes them inf / NaN and the flag false, exactly where the
// reference returns early.  Same expressions, same order, same rounding as tri_test (flx_trace.cuh).
FLX_DEV bool tri_test_flat(V3 v0, V3 s1, V3 s2, V3 o, V3 d, float &t, float &u, float &v)
{
    const V3 pvec = cross3(d, s2);
    const float det = dot3(s1, pvec);
    const float iDet = 1.0f / det;
    const V3 tvec = o - v0;
    u = dot3(tvec, pvec) * iDet;
    const V3 qvec = cross3(tvec, s1);
    v = dot3(d, qvec) * iDet;
    t = dot3(s2, qvec) * iDet;
    return !(fabsf(det) < 1e-12f) && !(u < 0.0f || u > 1.0f) && !(v < 0.0f || u + v > 1.0f) && !(t < 0.0f);
}

// lane states in `cur`: >= 0 inner node (TNode index); < 0 leaf cursor ~(TTri index), always > TG_DONE because the repack
// keeps the TTri count below 0x7ffffff0; TG_DONE ray finished, result not yet written; TG_IDLE no ray
#define TG_IDLE ((int)0x80000000)
#define TG_DONE ((int)0x80000001)

template <bool ANYHIT, class COUNT, int MIN_BLOCKS, int MODE = TRACE_WF>
__global__ void __launch_bounds__(FLX_TRACE_BLOCK, MIN_BLOCKS) k_trace_greedy(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const BvhView bvh,
                                                                              const flx_Triangle *tris160, uint32_t *fetchCounter, const int threshold, const int fetchChunk,
                                                                              const int innerBias, unsigned long long *countTotals, const MkView mk)
{
    static_assert(MODE == TRACE_WF || (MODE == TRACE_MK_NEXT && !ANYHIT) || (MODE == TRACE_MK_NEE && ANYHIT), "microkernel modes: closest hit for nextVertex, any hit for the light samples");
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned lanesBelow = (1u << lane) - 1u;
    const uint32_t *queue = MODE == TRACE_MK_NEE ? mk.rayQueue : fr.queues[ANYHIT ? Q_SHADOW : Q_EXT];
    const uint32_t count = MODE == TRACE_MK_NEXT ? mk.limit : (MODE == TRACE_MK_NEE ? *mk.rayCount : *counter_ptr(fr.counters, ANYHIT ? Q_SHADOW : Q_EXT));
    const Tasks &t = fr.tasks;
    const bool lightTest = ANYHIT ? (prm.useAreaLight != 0) : (prm.sampleImpl && prm.useAreaLight);

    uint32_t gid = 0;
    V3 o = v3(0.0f), d = v3(0.0f), idir = v3(0.0f);
    float tbest = 0.0f, ub = 0.0f, vb = 0.0f;
    int tri = -1, cur = TG_IDLE, sp = 0;
    bool occluded = false;
    int lstack[FLX_STACK_DEPTH];
    COUNT cnt;
    unsigned raysDone = 0;
    bool drain = false;                                                     // warp-uniform: the queue has nothing left to hand out
    uint32_t chunkNext = 0, chunkEnd = 0, chunkSize = (uint32_t)fetchChunk; // warp-uniform

    while (true)
    {
        // ---- write back finished rays (all lanes that finished since the last round do this together)
        if (cur == TG_DONE)
        {
            cur = TG_IDLE;
            raysDone++;
            if (ANYHIT && MODE == TRACE_MK_NEE)
                mk.scratch.setu_cs(MK_X_BLOCKED0 + (int)(gid & 1u), gid >> 1, occluded ? 1u : 0u);
            else if (ANYHIT)
                t.setu_cs(FLX_S_SHADOW_BLOCKED, gid, occluded ? 1u : 0u);
            else
            {
                V3 P = v3(0.0f), N = v3(0.0f);
                float tu = 0.0f, tv = 0.0f;
                int matId = -1, lightHit = 0;
                if (tri >= 0)
                {
                    P = o + tbest * d;
                    hit_attributes(bvh, tri, ub, vb, N, tu, tv, matId);
                }
                if (lightTest && light_quad(prm.areaLight, o, d, tbest)) // wf_extrays.cl:29
                {
                    lightHit = 1;
                    P = o + tbest * d;
                    N = v3(prm.areaLight.N);
                    tri = 0;
                    matId = 0;
                }
                t.setu_cs(FLX_S_PATH_LEN, gid, t.u_cs(FLX_S_PATH_LEN, gid) + 1u);
                t.setv_cs(FLX_S_P, gid, P);
                t.setv_cs(FLX_S_N, gid, N);
                t.setf_cs(FLX_S_UV, gid, tu);
                t.setf_cs(FLX_S_UV + 1, gid, tv);
                t.setf_cs(FLX_S_HIT_T, gid, tbest);
                t.setu_cs(FLX_S_HIT_I, gid, (uint32_t)tri);
                t.setu_cs(FLX_S_AREA_LIGHT_HIT, gid, (uint32_t)lightHit);
                t.setu_cs(FLX_S_MAT_ID, gid, (uint32_t)matId);
            }
        }

        // ---- idle lanes take the next rays of the queue.  A warp reserves the queue in chunks (one atomic per chunk, not per
        //      refill: every warp of the grid hits the same counter word and same-address atomics serialise in L2) and hands
        //      the chunk out locally; chunks shrink to 32 near the end of the queue to keep the tail balanced.
        const bool need = cur == TG_IDLE && !drain;
        const unsigned needMask = __ballot_sync(FULL, need);
        if (needMask)
        {
            const uint32_t needCount = (uint32_t)__popc(needMask);
            const uint32_t avail = chunkEnd - chunkNext; // warp-uniform
            uint32_t base = chunkNext, fresh = 0;
            if (avail < needCount)
            {
                if (lane == 0)
                    fresh = atomicAdd(fetchCounter, chunkSize);
                fresh = __shfl_sync(FULL, fresh, 0);
            }
            const uint32_t rank = (uint32_t)__popc(needMask & lanesBelow);
            uint32_t idx = base + rank;
            if (avail < needCount)
            {
                if (rank >= avail)
                    idx = fresh + (rank - avail);
                chunkNext = fresh + (needCount - avail);
                chunkEnd = fresh + chunkSize;
                if (fresh + chunkSize > count - count / 4u)
                    chunkSize = 32u;
            }
            else
                chunkNext += needCount;
            bool beyond = false;
            if (need)
            {
                if (idx < count && (MODE != TRACE_MK_NEXT || t.u_cs(FLX_S_PHASE, idx) == (uint32_t)MK_RT_NEXT_VERTEX))
                {
                    bool quadFirst = ANYHIT && lightTest; // the light quad is tested first and blocks (wf_shadowrays.cl:29-31)
                    if (MODE == TRACE_MK_NEE)
                    {
                        gid = __ldcs(queue + idx); // 2 * path + which
                        const uint32_t path = gid >> 1, which = gid & 1u;
                        o = mk.scratch.v_cs(MK_X_ORIG, path);
                        d = mk.scratch.v_cs(which ? MK_X_DIR1 : MK_X_DIR0, path);
                        // env-map sample: 2 * worldRadius, light quad blocks (mk_sample_bsdf.cl:82-90); area-light sample: its own
                        // length, no quad test (mk_sample_bsdf.cl:114-120)
                        tbest = which ? mk.scratch.f_cs(MK_X_LEN1, path) : 2.0f * prm.worldRadius;
                        quadFirst = lightTest && which == 0u;
                    }
                    else
                    {
                        gid = MODE == TRACE_MK_NEXT ? idx : __ldcs(queue + idx);
                        o = t.v_cs(ANYHIT ? FLX_S_SHADOW_ORIG : FLX_S_ORIG, gid);
                        d = t.v_cs(ANYHIT ? FLX_S_SHADOW_DIR : FLX_S_DIR, gid);
                        tbest = ANYHIT ? t.f_cs(FLX_S_SHADOW_RAY_LEN, gid) : 3.402823466e+38f;
                    }
                    idir = v3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    ub = vb = 0.0f;
                    tri = -1;
                    occluded = false;
                    cur = bvh.rootRef;
                    sp = 0;
                    if (cur < 0)
                        cnt.leaf(); // a scene whose root is a leaf
                    if (quadFirst)
                    {
                        float tl = tbest;
                        if (light_quad(prm.areaLight, o, d, tl))
                        {
                            occluded = true;
                            cur = TG_DONE;
                        }
                    }
                }
                else if (idx >= count)
                    beyond = true;
            }
            drain = __any_sync(FULL, beyond); // nothing left to fetch: run the remaining rays to the end
        }
        if (__ballot_sync(FULL, cur != TG_IDLE) == 0u && (MODE != TRACE_MK_NEXT || drain))
            break; // queue drained and every lane idle (TRACE_MK_NEXT: a round may draw only paths in other phases)

        // ---- traverse: one step of the majority kind per iteration, until too few lanes hold a ray
        while (true)
        {
            const bool sI = cur >= 0;
            const bool sT = cur < 0 && cur > TG_DONE;
            const int nI = __popc(__ballot_sync(FULL, sI)), nT = __popc(__ballot_sync(FULL, sT));
            if (nI + nT == 0 || (!drain && nI + nT < threshold))
                break;
            if (nT == 0 || (nI > 0 && nI + innerBias >= nT))
            {
                if (sI)
                {
                    cnt.inner();
                    const float4 *n = bvh.nodes + 4 * (size_t)cur;
                    const F8 h0 = ldg256_hint<FLX_HINT_NODE>(n), h1 = ldg256_hint<FLX_HINT_NODE>(n + 2);
                    const int cl = __float_as_int(h1.v[4]), cr = __float_as_int(h1.v[5]);
                    float ln, rn;
                    const bool lh = box_test(h0.v[0], h0.v[1], h0.v[2], h0.v[3], h0.v[4], h0.v[5], o, idir, tbest, ln);
                    const bool rh = box_test(h0.v[6], h0.v[7], h1.v[0], h1.v[1], h1.v[2], h1.v[3], o, idir, tbest, rn);
                    if (lh && rh)
                    {
                        const bool swap = rn < ln; // right child closer -> first (bvh.cl:292); ties keep left first
                        lstack[sp++] = swap ? cl : cr;
                        cur = swap ? cr : cl;
                    }
                    else if (lh)
                        cur = cl;
                    else if (rh)
                        cur = cr;
                    else if (sp > 0)
                        cur = lstack[--sp];
                    else
                        cur = TG_DONE;
                    if (cur < 0 && cur > TG_DONE)
                        cnt.leaf();
                }
            }
            else if (sT)
            {
                const float4 *p = bvh.tris + 4 * (size_t)(~cur);
                const F8 h0 = ldg256_hint<FLX_HINT_TRI>(p), h1 = ldg256_hint<FLX_HINT_TRI>(p + 2);
                const int tag = __float_as_int(h0.v[3]);
                float tt, uu, vv;
                cnt.tri();
                const bool hit = tri_test_flat(v3(h0.v[0], h0.v[1], h0.v[2]), v3(h0.v[4], h0.v[5], h0.v[6]), v3(h1.v[0], h1.v[1], h1.v[2]), o, d, tt, uu, vv) && tt > 0.0f && tt < tbest;
                if (ANYHIT)
                {
                    if (hit)
                        occluded = true;
                }
                else if (hit)
                {
                    cnt.hit();
                    tri = tag & 0x7fffffff;
                    tbest = tt;
                    ub = uu;
                    vb = vv;
                }
                if ((ANYHIT && hit) || (tag < 0 && sp == 0))
                {
                    cnt.leafEnd();
                    cur = TG_DONE;
                }
                else if (tag < 0) // last triangle of the leaf: on to the next node
                {
                    cnt.leafEnd();
                    cur = lstack[--sp];
                    if (cur < 0)
                        cnt.leaf();
                }
                else
                    cur -= 1; // ~(index + 1)
            }
        }
    }
    flush_counts(cnt, countTotals, raysDone);
}
