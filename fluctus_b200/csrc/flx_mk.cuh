// flx_mk.cuh -- the reference's second integrator, the LuxRender-style "microkernel" path tracer (one path per pixel,
// a phase word per path), as CUDA kernels for sm_100a.  It is the integrator Tracer::renderSingle uses for final frames
// (src/tracer.cpp:95-169: "only MK can guarantee given spp for every pixel") and the non-wavefront branch of
// Tracer::update (src/tracer.cpp:267-299).  SURVEY 8(f-4).
//
//   k_mk_reset          (reference: src/mk_reset.cl:4-43)
//   k_mk_raygen         (reference: src/mk_raygen.cl:5-63)
//   nextVertex          (reference: src/mk_next_vertex.cl:7-123)  = k_trace_persistent<closest, TRACE_MK_NEXT> + k_mk_next_vertex_logic
//   sampleBsdf          (reference: src/mk_sample_bsdf.cl:11-197) = k_mk_nee_prepare + k_trace_persistent<any, TRACE_MK_NEE> + k_mk_shade<type> per BSDF type
//   k_mk_splat          (reference: src/mk_splat.cl:5-41)
//   k_mk_splat_preview  (reference: src/mk_splat_preview.cl:5-25)
//
// What differs from the reference is the organisation, not the arithmetic.  The reference traces its rays inline, one
// thread per path (its own comment at src/mk_sample_bsdf.cl:86: "BAD! Collect all shadow ray casts together"); here both
// traversals go through the persistent-threads kernel of flx_trace_persistent.cuh: nextVertex hands it the paths whose phase
// is MK_RT_NEXT_VERTEX, and sampleBsdf is split in three -- light sampling (writes up to two candidate shadow rays per path
// to a scratch SoA and appends them to a ray list), one any-hit launch over that list, then shading.  The path state after
// each ABI call is bit-identical to the reference's kernel (the scratch is private to the call).
#pragma once

#include "flx_kernels.cuh"

// PathPhase, src/geom.h:184-193
enum { MK_RT_NEXT_VERTEX = 0, MK_SAMPLE_BSDF = 1, MK_SAMPLE_LIGHT_IMPL = 2, MK_HIT_NOTHING = 3, MK_SPLAT_SAMPLE = 4, MK_GENERATE_CAMERA_RAY = 5, MK_DONE = 6 };

// scratch SoA of the sampleBsdf call: slot s of path g at scratch[s * numTasks + g]
enum {
    MK_X_ORIG = 0,     // shadow-ray origin, shared by both samples (mk_sample_bsdf.cl:55)
    MK_X_DIR0 = 3,     // env-map sample: direction (length is 2 * worldRadius)
    MK_X_DIR1 = 6,     // area-light sample: direction
    MK_X_LEN1 = 9,     //                    distance to the sampled point
    MK_X_PDF0 = 10,    // env-map sample: directPdfW
    MK_X_SEED = 11,    // RNG state after the light samples
    MK_X_BLOCKED0 = 12, MK_X_BLOCKED1 = 13, // results of the any-hit launch
    MK_X_SLOTS = 14
};

// Shading lists: like the wavefront integrator's material queues (wf_logic.cl:322-372), so that a warp of the shading kernel sees
// one BSDF.  The reference's sampleBsdf evaluates whatever material each path hit in one kernel; the result per path is the same.
enum { MK_L_DIFFUSE = 0, MK_L_GLOSSY, MK_L_GGX_REFL, MK_L_GGX_REFR, MK_L_DELTA, MK_L_OTHER, MK_NUM_LISTS };
FLX_DEV int mk_list_of(int type)
{
    if (type == FLX_BXDF_DIFFUSE) return MK_L_DIFFUSE;
    if (type == FLX_BXDF_GLOSSY) return MK_L_GLOSSY;
    if (type == FLX_BXDF_GGX_ROUGH_REFLECTION) return MK_L_GGX_REFL;
    if (type == FLX_BXDF_GGX_ROUGH_DIELECTRIC) return MK_L_GGX_REFR;
    if (type == FLX_BXDF_IDEAL_REFLECTION || type == FLX_BXDF_IDEAL_DIELECTRIC) return MK_L_DELTA;
    return MK_L_OTHER; // emissive or unknown: the all-lobes kernel (bxdf.cl's switch falls through to black / white)
}

struct MkView
{
    Tasks scratch;
    uint32_t *rayQueue; // entries 2 * path + which
    uint32_t *rayCount;
    uint32_t *typeQueues; // MK_NUM_LISTS lists of numTasks path indices: the vertices to shade, by BSDF type
    uint32_t *typeCounts; // MK_NUM_LISTS counts
    uint32_t limit;     // min(width * height, numTasks): the paths the microkernels touch (e.g. mk_raygen.cl:9)
    flx_RenderStats64 *stats;
};

// CTA-wide sum of a per-thread count, one 64-bit atomic per CTA (reference: one atomic_inc per work-item on RenderStats,
// e.g. mk_next_vertex.cl:53, mk_sample_bsdf.cl:91; or per warp with -DNVIDIA)
FLX_DEV void mk_stat_add(unsigned long long *dst, uint32_t mine, uint32_t *s_part)
{
    const uint32_t w = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31) == 0)
        s_part[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x == 0)
    {
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < FLX_BLOCK / 32; i++)
            tot += s_part[i];
        if (tot)
            atomicAdd(dst, (unsigned long long)tot);
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ reset
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_reset(const __grid_constant__ Frame fr, const uint32_t limit)
{
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    if (gid >= limit)
        return;
    reinterpret_cast<float4 *>(fr.pixels)[gid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    reinterpret_cast<float4 *>(fr.denoiserNormal)[gid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    reinterpret_cast<float4 *>(fr.denoiserAlbedo)[gid] = make_float4(0.1f, 0.1f, 0.1f, 0.0f);
    fr.dirty[gid] = 1;
    const Tasks &t = fr.tasks;
    t.setu(FLX_S_PHASE, gid, (uint32_t)MK_GENERATE_CAMERA_RAY);
    t.setv(FLX_S_EI, gid, v3(0.0f));
    t.setv(FLX_S_T, gid, v3(1.0f));
    t.setu(FLX_S_PATH_LEN, gid, 0u);
    t.setu(FLX_S_LAST_SPECULAR, gid, 1u);
    t.setf(FLX_S_LAST_PDF_W, gid, 1.0f);
    t.setu(FLX_S_FIRST_DIFFUSE, gid, 0u);
    t.setu(FLX_S_SEED, gid, gid);
}

// ------------------------------------------------------------------------------------------------ camera rays
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_raygen(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const uint32_t limit)
{
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    if (gid >= limit)
        return;
    const Tasks &t = fr.tasks;
    if (t.u(FLX_S_PHASE, gid) != (uint32_t)MK_GENERATE_CAMERA_RAY)
        return;
    uint32_t seed = t.u(FLX_S_SEED, gid);
    uint32_t px, py; // path g renders local pixel g; tiled contexts map it to the full image like k_raygen
    local_pixel_to_xy(fr, prm.width, gid, px, py);
    float x = (float)px, y = (float)py;
    x += flx_rand(seed);
    y += flx_rand(seed);
    const float NDCx = x / (float)prm.width, NDCy = y / (float)prm.height;
    float SCRx = 2.0f * NDCx - 1.0f, SCRy = 2.0f * NDCy - 1.0f;
    SCRx *= (float)prm.width / (float)prm.height;
    SCRx *= fr.tanHalfFov;
    SCRy *= fr.tanHalfFov;
    const V3 camPos = v3(prm.camera.pos), camRight = v3(prm.camera.right), camUp = v3(prm.camera.up), camDir = v3(prm.camera.dir);
    V3 rayOrig = camPos;
    const V3 target = ((rayOrig + camRight * SCRx) + camUp * SCRy) + camDir;
    V3 rayDir = norm3(target - rayOrig);
    const V3 fp = camPos + rayDir * prm.camera.focalDist; // depth of field, mk_raygen.cl:49-53
    const float sqrt_r = sqrtf(flx_rand(seed));
    const float th = FLX_2PI_F * flx_rand(seed);
    rayOrig = rayOrig + lens_offset(fr, prm, camRight, camUp, sqrt_r, th);
    rayDir = norm3(fp - rayOrig);
    t.setv(FLX_S_ORIG, gid, rayOrig);
    t.setv(FLX_S_DIR, gid, rayDir);
    t.setu(FLX_S_SEED, gid, seed);
    t.setu(FLX_S_PHASE, gid, (uint32_t)MK_RT_NEXT_VERTEX);
}

// ------------------------------------------------------------------------------------------------ nextVertex, after the trace
// The traversal launch has written the hit record and pathLen + 1 for every path in phase MK_RT_NEXT_VERTEX (the same
// write-back as the wavefront extension stage); this kernel does the rest of mk_next_vertex.cl:47-122: ray statistics,
// implicit environment / area-light samples with MIS, and the phase change.
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_next_vertex_logic(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const SceneView sc,
                                                                    const MkView mk)
{
    __shared__ uint32_t s_part[FLX_BLOCK / 32];
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    const Tasks &t = fr.tasks;
    const bool live = gid < mk.limit && t.u(FLX_S_PHASE, gid) == (uint32_t)MK_RT_NEXT_VERTEX;
    uint32_t primary = 0, extension = 0;
    if (live)
    {
        const uint32_t len = t.u(FLX_S_PATH_LEN, gid); // already incremented
        (len == 1u ? primary : extension) = 1u;
        if (fr.denoiser && len == 1u) // mk_next_vertex.cl:60-70: first-hit normal (zero for a miss); path g owns pixel g, no atomics
        {
            float4 *dst = reinterpret_cast<float4 *>(fr.denoiserNormal) + gid;
            const float4 prev = *dst, add = denoiser_normal(prm, t.v(FLX_S_N, gid));
            *dst = make_float4(prev.x + add.x, prev.y + add.y, prev.z + add.z, prev.w + add.w);
        }
        const int hitI = (int)t.u(FLX_S_HIT_I, gid);
        const bool hitLight = t.u(FLX_S_AREA_LIGHT_HIT, gid) != 0u;
        uint32_t phase = (uint32_t)MK_SAMPLE_BSDF;
        if (hitI < 0 || hitLight)
        {
            const V3 rayOrig = t.v(FLX_S_ORIG, gid), rayDir = t.v(FLX_S_DIR, gid);
            const V3 T = t.v(FLX_S_T, gid), Ei = t.v(FLX_S_EI, gid);
            const bool lastSpecular = t.u(FLX_S_LAST_SPECULAR, gid) != 0u;
            const float lastPdfW = t.f(FLX_S_LAST_PDF_W, gid);
            V3 newEi;
            if (hitI < 0) // implicit environment-map sample, mk_next_vertex.cl:73-94
            {
                V3 bg = v3(0.0f);
                if (prm.useEnvMap && (len == 1u || prm.sampleImpl))
                    bg = env_eval_dir(sc, rayDir) * prm.envMapStrength;
                float weight = 1.0f;
                if (prm.sampleImpl && prm.sampleExpl && prm.useEnvMap && len > 1u && !lastSpecular)
                {
                    const float lightPickProb = 1.0f;
                    const float directPdfW = env_pdf(sc, rayDir);
                    weight = (lastPdfW * lightPickProb) / (lastPdfW * lightPickProb + directPdfW);
                }
                newEi = Ei + (weight * T) * bg;
            }
            else // implicit area-light sample, mk_next_vertex.cl:96-116
            {
                float misWeight = 1.0f;
                if (prm.sampleExpl && len > 1u && !lastSpecular)
                {
                    const V3 hP = t.v(FLX_S_P, gid), hN = t.v(FLX_S_N, gid);
                    const float directPdfA = 1.0f / (4.0f * prm.areaLight.size.x * prm.areaLight.size.y);
                    const float dist = len3(hP - rayOrig);
                    const float cosine = dot3(norm3(-rayDir), hN);
                    const float directPdfW = directPdfA * (dist * dist) / fabsf(cosine); // pdfAtoW, utils.cl:197-200
                    const float lightPickProb = 1.0f;
                    misWeight = lastPdfW / (lastPdfW + directPdfW * lightPickProb);
                }
                newEi = Ei + (T * misWeight) * v3(prm.areaLight.E);
            }
            t.setv(FLX_S_EI, gid, newEi);
            phase = (uint32_t)MK_SPLAT_SAMPLE;
        }
        t.setu(FLX_S_PHASE, gid, phase);
    }
    mk_stat_add(reinterpret_cast<unsigned long long *>(&mk.stats->primaryRays), primary, s_part);
    mk_stat_add(reinterpret_cast<unsigned long long *>(&mk.stats->extensionRays), extension, s_part);
}

// ------------------------------------------------------------------------------------------------ sampleBsdf, part 1: light samples
// what both halves of sampleBsdf derive from the path state before any random number is drawn (mk_sample_bsdf.cl:40-55)
struct MkVertex
{
    Surface s; // N already normal-mapped and flipped to the incoming side
    Mat mat;
    V3 rayDir, orig;
    bool backface, singular;
};
FLX_DEV MkVertex mk_load_vertex(const Tasks &t, uint32_t gid, const SceneView &sc)
{
    MkVertex v;
    v.rayDir = t.v(FLX_S_DIR, gid);
    v.s.P = t.v(FLX_S_P, gid);
    v.s.N = t.v(FLX_S_N, gid);
    v.s.u = t.f(FLX_S_UV, gid);
    v.s.v = t.f(FLX_S_UV + 1, gid);
    v.s.tri = (int)t.u(FLX_S_HIT_I, gid);
    v.mat = load_material(sc.materials, (int)t.u(FLX_S_MAT_ID, gid), sc.kdGamma);
    V3 N = shading_normal(v.s, v.mat, sc);
    v.backface = dot3(N, v.rayDir) > 0.0f;
    if (v.backface)
        N = N * -1.0f;
    v.s.N = N;
    v.orig = v.s.P - 1e-3f * v.rayDir; // avoid self-shadowing
    v.singular = (v.mat.type & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC)) != 0;
    return v;
}

__global__ void __launch_bounds__(FLX_BLOCK) k_mk_nee_prepare(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const SceneView sc,
                                                              const MkView mk)
{
    __shared__ uint32_t s_part[FLX_BLOCK / 32];
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    const Tasks &t = fr.tasks;
    const Tasks &x = mk.scratch;
    const bool live = gid < mk.limit && t.u(FLX_S_PHASE, gid) == (uint32_t)MK_SAMPLE_BSDF;
    uint32_t shadowRays = 0;
    bool push0 = false, push1 = false;
    int list = -1;
    if (live)
    {
        uint32_t seed = t.u(FLX_S_SEED, gid);
        const MkVertex v = mk_load_vertex(t, gid, sc);
        list = mk_list_of(v.mat.type);
        if (fr.denoiser && !v.singular && t.u(FLX_S_FIRST_DIFFUSE, gid) == 0u) // mk_sample_bsdf.cl:56-66
        {
            t.setu(FLX_S_FIRST_DIFFUSE, gid, 1u);
            const V3 albedo = mat_float3(v.mat.Kd, v.s.u, v.s.v, v.mat.map_Kd, sc); // not gamma-corrected
            float4 *dst = reinterpret_cast<float4 *>(fr.denoiserAlbedo) + gid;
            const float4 prev = *dst;
            *dst = make_float4(prev.x + albedo.x, prev.y + albedo.y, prev.z + albedo.z, prev.w + 1.0f);
        }
        if (prm.sampleExpl && !v.singular)
        {
            x.setv(MK_X_ORIG, gid, v.orig);
            if (prm.useEnvMap) // mk_sample_bsdf.cl:73-91
            {
                V3 L;
                float directPdfW = 0.0f;
                env_sample_alias(sc, flx_rand(seed), L, directPdfW);
                L = norm3(L);
                x.setv(MK_X_DIR0, gid, L);
                x.setf(MK_X_PDF0, gid, directPdfW);
                shadowRays++;
                push0 = directPdfW != 0.0f; // the reference traces regardless and then ignores the answer (mk_sample_bsdf.cl:94)
            }
            if (prm.useAreaLight) // mk_sample_bsdf.cl:114-121; sampleAreaLight utils.cl:226-234
            {
                const flx_AreaLight &A = prm.areaLight;
                V3 posL = v3(A.pos);
                const float r1 = 2.0f * flx_rand(seed) - 1.0f;
                const float r2 = 2.0f * flx_rand(seed) - 1.0f;
                posL = posL + (r1 * A.size.x) * v3(A.right);
                posL = posL + (r2 * A.size.y) * v3(A.up);
                V3 L = posL - v.orig;
                const float lenL = len3(L);
                L = norm3(L);
                x.setv(MK_X_DIR1, gid, L);
                x.setf(MK_X_LEN1, gid, lenL);
                shadowRays++;
                push1 = fmaxf(dot3(v3(A.N), -L), 0.0f) > 0.0f; // same: a sample on the light's back side is traced but unused (:125)
            }
        }
        x.setu(MK_X_SEED, gid, seed);
    }
    // append the rays that matter to the list: one atomic per warp
    const uint32_t n = (push0 ? 1u : 0u) + (push1 ? 1u : 0u);
    const int lane = threadIdx.x & 31;
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o)
            incl += y;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    uint32_t base = 0;
    if (lane == 31 && total)
        base = atomicAdd(mk.rayCount, total);
    base = __shfl_sync(0xffffffffu, base, 31);
    uint32_t slot = base + incl - n;
    if (push0)
        mk.rayQueue[slot++] = 2u * gid;
    if (push1)
        mk.rayQueue[slot] = 2u * gid + 1u;
    // ... and every vertex to the shading list of its BSDF type: one atomic per warp per type present
#pragma unroll
    for (int k = 0; k < MK_NUM_LISTS; k++)
    {
        const unsigned m = __ballot_sync(0xffffffffu, list == k);
        if (m == 0u)
            continue;
        uint32_t lbase = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader)
            lbase = atomicAdd(mk.typeCounts + k, (uint32_t)__popc(m));
        lbase = __shfl_sync(0xffffffffu, lbase, leader);
        if (list == k)
            mk.typeQueues[(size_t)k * fr.numTasks + lbase + __popc(m & ((1u << lane) - 1u))] = gid;
    }
    mk_stat_add(reinterpret_cast<unsigned long long *>(&mk.stats->shadowRays), shadowRays, s_part);
}

// ------------------------------------------------------------------------------------------------ sampleBsdf, part 2: shading
template <int ALL> // the lobes compiled in: one shading list's BSDF type(s), like the wavefront material kernels
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_shade(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const SceneView sc, const MkView mk,
                                                        const int list)
{
    const Tasks &t = fr.tasks;
    const Tasks &x = mk.scratch;
    const uint32_t count = mk.typeCounts[list];
    const uint32_t *queue = mk.typeQueues + (size_t)list * fr.numTasks;
    for (uint32_t idx = blockIdx.x * FLX_BLOCK + threadIdx.x; idx < count; idx += gridDim.x * FLX_BLOCK)
    {
    const uint32_t gid = queue[idx];
    uint32_t seed = x.u(MK_X_SEED, gid);
    const MkVertex v = mk_load_vertex(t, gid, sc);
    const V3 T = t.v(FLX_S_T, gid);

    if (prm.sampleExpl && !v.singular) // next-event estimation, mk_sample_bsdf.cl:69-147
    {
        const float lightPickProb = 1.0f;
        if (prm.useEnvMap)
        {
            const float directPdfW = x.f(MK_X_PDF0, gid);
            if (directPdfW != 0.0f && x.u(MK_X_BLOCKED0, gid) == 0u)
            {
                const V3 L = x.v(MK_X_DIR0, gid);
                const V3 brdf = bxdf_eval<ALL>(v.s, v.mat, v.backface, sc, v.rayDir, L);
                const float cosTh = fmaxf(0.0f, dot3(L, v.s.N));
                const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf<ALL>(v.s, v.mat, v.backface, sc, v.rayDir, L));
                float weight = 1.0f;
                if (prm.sampleImpl)
                    weight = (directPdfW * lightPickProb) / (directPdfW * lightPickProb + bsdfPdfW);
                const V3 envMapLi = env_eval_dir(sc, L) * prm.envMapStrength;
                const V3 contrib = ((((brdf * T) * envMapLi) * weight) * cosTh) / (lightPickProb * directPdfW);
                t.setv(FLX_S_EI, gid, t.v(FLX_S_EI, gid) + contrib);
            }
        }
        if (prm.useAreaLight)
        {
            const V3 L = x.v(MK_X_DIR1, gid);
            const float cosLight = fmaxf(dot3(v3(prm.areaLight.N), -L), 0.0f); // only frontside hits count
            if (cosLight > 0.0f && x.u(MK_X_BLOCKED1, gid) == 0u)
            {
                const float lenL = x.f(MK_X_LEN1, gid);
                const float directPdfA = 1.0f / (4.0f * prm.areaLight.size.x * prm.areaLight.size.y);
                const V3 brdf = bxdf_eval<ALL>(v.s, v.mat, v.backface, sc, v.rayDir, L);
                const float cosTh = fmaxf(0.0f, dot3(L, v.s.N));
                const float directPdfW = directPdfA * (lenL * lenL) / fabsf(cosLight); // pdfAtoW
                const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf<ALL>(v.s, v.mat, v.backface, sc, v.rayDir, L));
                float weight = 1.0f;
                if (prm.sampleImpl)
                    weight = (directPdfW * lightPickProb) / (directPdfW * lightPickProb + bsdfPdfW);
                const V3 contrib = ((((brdf * T) * v3(prm.areaLight.E)) * weight) * cosTh) / (lightPickProb * directPdfW);
                t.setv(FLX_S_EI, gid, t.v(FLX_S_EI, gid) + contrib);
            }
        }
    }

    // path termination (Russian roulette), mk_sample_bsdf.cl:149-157
    float contProb = 1.0f;
    const uint32_t len = t.u(FLX_S_PATH_LEN, gid);
    bool terminate = (len - 1u >= prm.maxBounces);
    if (terminate && prm.useRoulette)
    {
        contProb = fminf(fmaxf(luminance3(T), 0.01f), 0.5f);
        terminate = (flx_rand(seed) > contProb);
    }

    // continuation ray, mk_sample_bsdf.cl:159-193.  pdfW starts at 0: the reference leaves it uninitialised when sampleGlossy
    // rejects a direction (src/glossy.cl:58-59); the path terminates then, so nothing downstream reads it
    float pdfW = 0.0f;
    V3 newDir = v3(0.0f);
    const V3 bsdf = bxdf_sample<ALL>(v.s, v.mat, v.backface, sc, v.rayDir, newDir, pdfW, seed);
    const float costh = dot3(v.s.N, norm3(newDir));
    pdfW *= contProb;
    if (pdfW == 0.0f || is_zero3(bsdf))
        terminate = true;
    const V3 newT = ((T * bsdf) * costh) / pdfW;
    const V3 orig = v.s.P + 1e-4f * newDir;
    t.setv(FLX_S_T, gid, newT);
    t.setv(FLX_S_ORIG, gid, orig);
    t.setv(FLX_S_DIR, gid, newDir);
    t.setf(FLX_S_LAST_PDF_W, gid, pdfW);
    t.setu(FLX_S_SEED, gid, seed);
    t.setu(FLX_S_LAST_SPECULAR, gid, v.singular ? 1u : 0u);
    t.setu(FLX_S_PHASE, gid, terminate ? (uint32_t)MK_SPLAT_SAMPLE : (uint32_t)MK_RT_NEXT_VERTEX);
    }
}

// ------------------------------------------------------------------------------------------------ splat
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_splat(const __grid_constant__ Frame fr, const MkView mk)
{
    __shared__ uint32_t s_part[FLX_BLOCK / 32];
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    const Tasks &t = fr.tasks;
    const bool live = gid < mk.limit && t.u(FLX_S_PHASE, gid) == (uint32_t)MK_SPLAT_SAMPLE;
    if (live)
    {
        const V3 Ei = t.v(FLX_S_EI, gid);
        float4 color = make_float4(Ei.x, Ei.y, Ei.z, 1.0f);
        float4 *px = reinterpret_cast<float4 *>(fr.pixels) + gid; // path g owns pixel g: no atomics (mk_splat.cl:20-24)
        const float4 prev = *px;
        if (prev.w > 0.0f)
            color = make_float4(color.x + prev.x, color.y + prev.y, color.z + prev.z, color.w + prev.w);
        *px = color;
        fr.dirty[gid] = 1;
        t.setv(FLX_S_EI, gid, v3(0.0f));
        t.setv(FLX_S_T, gid, v3(1.0f));
        t.setu(FLX_S_PATH_LEN, gid, 0u);
        t.setu(FLX_S_FIRST_DIFFUSE, gid, 0u);
        t.setu(FLX_S_PHASE, gid, (uint32_t)MK_GENERATE_CAMERA_RAY);
    }
    mk_stat_add(reinterpret_cast<unsigned long long *>(&mk.stats->samples), live ? 1u : 0u, s_part);
}

// interactive preview: every path splats what it has, alpha 0 forces an overwrite next time (mk_splat_preview.cl:5-25)
__global__ void __launch_bounds__(FLX_BLOCK) k_mk_splat_preview(const __grid_constant__ Frame fr, const uint32_t limit)
{
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    if (gid >= limit)
        return;
    const Tasks &t = fr.tasks;
    const V3 Ei = t.v(FLX_S_EI, gid);
    reinterpret_cast<float4 *>(fr.pixels)[gid] = make_float4(Ei.x, Ei.y, Ei.z, 0.0f);
    fr.dirty[gid] = 1;
    t.setv(FLX_S_EI, gid, v3(0.0f));
    t.setv(FLX_S_T, gid, v3(1.0f));
    t.setu(FLX_S_PATH_LEN, gid, 0u);
    t.setu(FLX_S_PHASE, gid, (uint32_t)MK_GENERATE_CAMERA_RAY);
}
