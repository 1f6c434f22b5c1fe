// flx_trace_persistent.cuh -- the production traversal kernels: persistent threads with dynamic ray fetch.
//
// Why: with one ray per thread (k_extrays/k_shadowrays in flx_kernels.cuh, kept as the simple variant) ncu shows 5.9 of
// 32 lanes active per issued instruction on Conference (profiles/old/r1_v1_extrays_raw.csv): incoherent rays differ widely in
// traversal length, so a warp is held by its longest ray.  Here a warp owns 32 ray SLOTS instead: whenever fewer than
// `threshold` lanes still hold a ray, the idle lanes take new rays from the queue with one warp-aggregated atomic
// (persistent-threads scheme after Aila & Laine 2009), and the traversal itself is organised "while-while": all lanes
// walk inner nodes until each has reached a leaf (or finished), then all intersect their leaves.
//
// Exactness: per ray the visiting order, box test and triangle test are those of flx_trace.cuh (i.e. the reference's),
// so results stay bit-identical; only WHICH lane traces a ray and WHEN changes.  No speculative traversal for closest
// hits: testing boxes against a stale hit distance can admit a leaf the reference culls, and on axis-aligned geometry a
// coplanar triangle there can win a tie the reference never sees.
//
// Memory path (what the ncu capture of the first persistent version showed: the L1TEX tag stage, one wavefront per
// lane per load instruction for divergent 16-byte loads, was the limiter at ~1.1 wavefronts/clk/SM):
//   * nodes and triangles are fetched with 256-bit loads (LDG.E.256, new on sm_100): 2 wavefronts per record, not 4 / 3, with L1
//     eviction hints (inner nodes evict-last, leaf triangles evict-first, hit attributes without allocation: flx_trace.cuh);
//   * the instruction stream (round 2, once the loads were no longer the only limiter): a lane's state IS its node reference (inner
//     node / leaf / TR_DONE / TR_IDLE) instead of flags the compiler packs and unpacks, and a lane takes 2 (closest hit) or 3 (any hit)
//     node steps per warp vote -- 10 % fewer issued instructions per launch, same results;
//   * TOP variant: the hottest part of the tree -- a treelet grown from the root by always expanding the node with the
//     largest box area, which repack_bvh lays out FIRST in the node array -- is staged once per CTA into shared memory
//     by the bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) and read with LDS.128.  On Conference a 2047-node
//     treelet (128 KB) serves 92 % of all inner-node visits (measured with the instrumented oracle).  One persistent
//     CTA per SM owns the staged treelet.
#pragma once

#include "flx_kernels.cuh"
#include "flx_mk.cuh"

FLX_DEV uint32_t smem_addr(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

// Stage `bytes` (multiple of 16) from global to shared memory with the bulk-copy engine; all threads of the CTA call this.
FLX_DEV void stage_bulk(void *dstShared, const void *srcGlobal, uint32_t bytes, unsigned long long *mbar)
{
    const uint32_t bar = smem_addr(mbar);
    if (threadIdx.x == 0)
    {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0 && bytes > 0)
    {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        const uint32_t chunk = 32768u;
        for (uint32_t off = 0; off < bytes; off += chunk)
        {
            const uint32_t n = min(chunk, bytes - off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dstShared) + off),
                         "l"(reinterpret_cast<const char *>(srcGlobal) + off), "r"(n), "r"(bar)
                         : "memory");
        }
    }
    if (bytes > 0)
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
    }
}

// MODE selects where the rays come from and where the result goes:
//   TRACE_WF       the wavefront stages: extension / shadow queue of path indices, rays and results in the path state;
//   TRACE_MK_NEXT  microkernel nextVertex (src/mk_next_vertex.cl:22-40): no queue -- every path g < limit whose phase is
//                  MK_RT_NEXT_VERTEX is traced; a lane that draws a path in another phase simply asks again;
//   TRACE_MK_NEE   microkernel next-event rays (src/mk_sample_bsdf.cl:87-91, 120-121): the ray list and the rays live in the
//                  MkView scratch (flx_mk.cuh), two candidate rays per path, entry = 2 * path + which.
enum { TRACE_WF = 0, TRACE_MK_NEXT = 1, TRACE_MK_NEE = 2 };

template <bool ANYHIT, class COUNT, int BLOCK, bool TOP, int MIN_BLOCKS, int SDEPTH, int MODE = TRACE_WF>
__global__ void __launch_bounds__(BLOCK, MIN_BLOCKS) k_trace_persistent(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm,
                                                                         const BvhView bvh, const flx_Triangle *tris160, uint32_t *fetchCounter,
                                                                         const int threshold, const int innerMin, const int fetchChunk, const int topCount, unsigned long long *countTotals,
                                                                         const MkView mk)
{
    static_assert(MODE == TRACE_WF || (MODE == TRACE_MK_NEXT && !ANYHIT) || (MODE == TRACE_MK_NEE && ANYHIT), "microkernel modes: closest hit for nextVertex, any hit for the light samples");
    constexpr unsigned FULL = 0xffffffffu;
    extern __shared__ __align__(128) unsigned char dynSmem[];
    __shared__ unsigned long long stageBar;
    const float4 *topNodes = reinterpret_cast<const float4 *>(dynSmem);
    if (TOP)
        stage_bulk(dynSmem, bvh.nodes, (uint32_t)topCount * 64u, &stageBar);
    const int lane = threadIdx.x & 31;
    const unsigned lanesBelow = (1u << lane) - 1u;
    const uint32_t *queue = MODE == TRACE_MK_NEE ? mk.rayQueue : fr.queues[ANYHIT ? Q_SHADOW : Q_EXT];
    const uint32_t count = MODE == TRACE_MK_NEXT ? mk.limit : (MODE == TRACE_MK_NEE ? *mk.rayCount : *counter_ptr(fr.counters, ANYHIT ? Q_SHADOW : Q_EXT));
    const Tasks &t = fr.tasks;
    const bool lightTest = ANYHIT ? (prm.useAreaLight != 0) : (prm.sampleImpl && prm.useAreaLight);

    // A lane's state is its node reference: an inner node (0 <= cur < TR_DONE), a leaf (cur < 0), TR_DONE = ray finished and its result not
    // yet written, TR_IDLE = no ray.  (Separate flags cost a dozen byte-permute instructions per node step: the compiler packs bools.)
    constexpr int TR_IDLE = 0x7fffffff, TR_DONE = 0x7ffffffe;
    bool exhausted = false;
    uint32_t gid = 0;
    V3 o = v3(0.0f), d = v3(0.0f), idir = v3(0.0f);
    float tbest = 0.0f, ub = 0.0f, vb = 0.0f;
    int tri = -1, cur = TR_IDLE, sp = 0;
    bool occluded = false;
    // Traversal stack: the first SDEPTH levels live in shared memory, laid out [level][thread] so that every lane always
    // hits its own bank whatever its depth (a push/pop is ONE wavefront); deeper levels -- beyond any tree the reference's
    // builders make for the shipped scenes -- spill to local memory.  With SDEPTH = 0 the whole stack is local memory, where
    // a push/pop by lanes at different depths touches up to 32 lines of L1, i.e. costs as much as a divergent node fetch.
    // SDEPTH = -1: the whole stack in local memory, but the most recent entry stays in a register (`topReg`): an entry that is popped
    // before the next push -- the far child of a node whose near child led straight to a leaf -- never touches memory.
    static_assert(!(TOP && SDEPTH > 0), "the treelet and the stack do not share dynamic shared memory");
    constexpr int SD = SDEPTH > 0 ? SDEPTH : 0;
    constexpr bool TOPREG = SDEPTH < 0;
    constexpr int NO_ENTRY = 0x7fffffff; // not a node reference (inner >= 0 and < 2^31 - 1, leaf < 0)
    int lstack[FLX_STACK_DEPTH - SD];
    int topReg = NO_ENTRY;
    int *const sstack = reinterpret_cast<int *>(dynSmem) + threadIdx.x;
#define FLX_PUSH(v)                                                                                                                                            \
    do                                                                                                                                                         \
    {                                                                                                                                                          \
        if (TOPREG)                                                                                                                                            \
        {                                                                                                                                                      \
            if (topReg != NO_ENTRY)                                                                                                                            \
                lstack[sp++] = topReg;                                                                                                                         \
            topReg = (v);                                                                                                                                      \
        }                                                                                                                                                      \
        else                                                                                                                                                   \
        {                                                                                                                                                      \
            if (SD == 0 || sp >= SD)                                                                                                                           \
                lstack[sp - SD] = (v);                                                                                                                         \
            else                                                                                                                                               \
                sstack[sp * BLOCK] = (v);                                                                                                                      \
            sp++;                                                                                                                                              \
        }                                                                                                                                                      \
    } while (0)
#define FLX_POP(dst)                                                                                                                                           \
    do                                                                                                                                                         \
    {                                                                                                                                                          \
        if (TOPREG && topReg != NO_ENTRY)                                                                                                                      \
        {                                                                                                                                                      \
            (dst) = topReg;                                                                                                                                    \
            topReg = NO_ENTRY;                                                                                                                                 \
        }                                                                                                                                                      \
        else                                                                                                                                                   \
        {                                                                                                                                                      \
            --sp;                                                                                                                                              \
            (dst) = (SD == 0 || sp >= SD) ? lstack[sp - SD] : sstack[sp * BLOCK];                                                                              \
        }                                                                                                                                                      \
    } while (0)
#define FLX_STACK_EMPTY (sp == 0 && (!TOPREG || topReg == NO_ENTRY))
    COUNT cnt;
    unsigned raysDone = 0;
    uint32_t chunkNext = 0, chunkEnd = 0, chunkSize = (uint32_t)fetchChunk; // warp-uniform

    while (true)
    {
        // ---- write back finished rays (all lanes that finished since the last round do this together)
        if (cur == TR_DONE)
        {
            cur = TR_IDLE;
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

        // ---- idle lanes take the next rays of the queue.  A warp reserves the queue in chunks (one atomic per chunk, not
        //      per refill: every warp of the grid hits the same counter word and same-address atomics serialise in L2)
        //      and hands the chunk out locally; chunks shrink to 32 near the end of the queue to keep the tail balanced.
        const bool need = cur == TR_IDLE && !exhausted;
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
                    topReg = NO_ENTRY;
                    if (quadFirst)
                    {
                        float tl = tbest;
                        if (light_quad(prm.areaLight, o, d, tl))
                        {
                            occluded = true;
                            cur = TR_DONE;
                        }
                    }
                }
                else if (idx >= count)
                    exhausted = true;
            }
        }
        const bool drain = __any_sync(FULL, exhausted); // nothing left to fetch: run the remaining rays to the end
        if (__ballot_sync(FULL, cur != TR_IDLE) == 0u && (MODE != TRACE_MK_NEXT || drain))
            break; // queue drained and every lane idle (TRACE_MK_NEXT: a round may draw only paths in other phases)

        // ---- traverse until too few lanes hold a ray
        while (true)
        {
            // (1) inner nodes: every lane that is at an inner node steps, until fewer than innerMin lanes still are
            //     (innerMin = 1: until every lane has reached a leaf or finished its ray)
            while (true)
            {
                const bool atInner = (unsigned)cur < (unsigned)TR_DONE;
                const unsigned innerMask = __ballot_sync(FULL, atInner);
                if (innerMask == 0u)
                    break;
                if (__popc(innerMask) < innerMin && __ballot_sync(FULL, cur < 0) != 0u)
                    break; // few stragglers and some lane has a leaf to intersect: switch phase (never with nothing to do)
                if (!atInner)
                    continue;
                // Several node steps per vote: a lane that is still at an inner node after its step takes the next one without asking the
                // warp again (lanes that reached a leaf meanwhile just wait: the result cannot change, only who idles when).  The two votes
                // of the loop header were 9 % of all issued instructions at 32 lanes; measured (Conference, kernels alone): closest hit
                // 0.702 -> 0.670 ms with 2 steps (0.681 with 3), any hit 0.345 -> 0.333 with 2, 0.329 with 3.
#ifndef FLX_INNER_STEPS_CLOSEST
#define FLX_INNER_STEPS_CLOSEST 2
#endif
#ifndef FLX_INNER_STEPS_ANY
#define FLX_INNER_STEPS_ANY 3
#endif
                constexpr int INNER_STEPS = ANYHIT ? FLX_INNER_STEPS_ANY : FLX_INNER_STEPS_CLOSEST;
#pragma unroll
                for (int rep = 0; rep < INNER_STEPS; rep++)
                {
                    if (rep > 0 && !((unsigned)cur < (unsigned)TR_DONE))
                        break;
                    cnt.inner();
                    float4 q0, q1, q2;
                    int4 q3;
                    if (TOP && cur < topCount)
                    {
                        const float4 *n = topNodes + 4 * cur;
                        q0 = n[0];
                        q1 = n[1];
                        q2 = n[2];
                        q3 = *reinterpret_cast<const int4 *>(n + 3);
                    }
                    else
                    {
                        const float4 *n = bvh.nodes + 4 * (size_t)cur;
                        const F8 h0 = ldg256_hint<FLX_HINT_NODE>(n), h1 = ldg256_hint<FLX_HINT_NODE>(n + 2);
                        q0 = make_float4(h0.v[0], h0.v[1], h0.v[2], h0.v[3]);
                        q1 = make_float4(h0.v[4], h0.v[5], h0.v[6], h0.v[7]);
                        q2 = make_float4(h1.v[0], h1.v[1], h1.v[2], h1.v[3]);
                        q3 = make_int4(__float_as_int(h1.v[4]), __float_as_int(h1.v[5]), 0, 0);
                    }
                    float ln, rn;
                    const bool lh = box_test(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, o, idir, tbest, ln);
                    const bool rh = box_test(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, o, idir, tbest, rn);
                    if (lh && rh)
                    {
                        // right child closer -> first (bvh.cl:292); ties keep left first.  (Any-hit rays could visit in any order; always
                        // taking the left child first was measured slower, 0.344 -> 0.364 ms: near-first finds occluders sooner.  So was
                        // prefetching both children while the box tests run.  Both knobs are gone from this loop: profiles/r2_knob_sweep.txt.)
                        const bool swap = rn < ln;
                        FLX_PUSH(swap ? q3.x : q3.y);
                        cur = swap ? q3.y : q3.x;
                    }
                    else if (lh)
                        cur = q3.x;
                    else if (rh)
                        cur = q3.y;
                    else if (!FLX_STACK_EMPTY)
                        FLX_POP(cur);
                    else
                    {
                        cur = TR_DONE;
                    }
                }
            }
            // (2) one leaf
            if (cur < 0)
            {
                cnt.leaf();
                const float4 *p = bvh.tris + 4 * (size_t)(~cur);
                float tmin = 3.402823466e+38f, umin = 0.0f, vmin = 0.0f;
                int imin = -1;
                while (true)
                {
                    const F8 h0 = ldg256_hint<FLX_HINT_TRI>(p), h1 = ldg256_hint<FLX_HINT_TRI>(p + 2);
                    const int tag = __float_as_int(h0.v[3]);
                    float tt, uu, vv;
                    cnt.tri();
                    if (tri_test(v3(h0.v[0], h0.v[1], h0.v[2]), v3(h0.v[4], h0.v[5], h0.v[6]), v3(h1.v[0], h1.v[1], h1.v[2]), o, d, tt, uu, vv))
                    {
                        if (ANYHIT)
                        {
                            if (tt > 0.0f && tt < tbest)
                            {
                                occluded = true;
                                break;
                            }
                        }
                        else if (tt > 0.0f && tt < tmin)
                        {
                            imin = tag & 0x7fffffff;
                            tmin = tt;
                            umin = uu;
                            vmin = vv;
                        }
                    }
                    if (tag < 0)
                        break;
                    p += 4;
                }
                if (!ANYHIT && imin != -1 && tmin < tbest)
                {
                    cnt.update();
                    tri = imin;
                    tbest = tmin;
                    ub = umin;
                    vb = vmin;
                }
                if ((ANYHIT && occluded) || FLX_STACK_EMPTY)
                {
                    cur = TR_DONE;
                }
                else
                    FLX_POP(cur);
            }
            const unsigned still = __ballot_sync(FULL, (unsigned)(cur - TR_DONE) > 1u); // lanes with a ray under way
            if (still == 0u)
                break;
            if (!drain && __popc(still) < threshold)
                break;
        }
    }
    flush_counts(cnt, countTotals, raysDone);
#undef FLX_PUSH
#undef FLX_POP
#undef FLX_STACK_EMPTY
}
