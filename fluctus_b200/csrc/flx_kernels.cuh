// flx_kernels.cuh -- the stages of one wavefront iteration as CUDA kernels for sm_100a:
//   k_reset        (reference: src/wf_reset.cl:5-66)
//   k_raygen       (reference: src/wf_raygen.cl:4-97)
//   k_extrays      (reference: src/wf_extrays.cl:5-36)
//   k_shadowrays   (reference: src/wf_shadowrays.cl:6-37)
//   k_logic        (reference: src/wf_logic.cl:14-314 and the three enqueue strategies 322-519)
//   k_material<M>  (reference: src/wf_mat_*.cl, one instantiation per material queue)
//   k_end_iteration (what the reference's host does between iterations: tracer.cpp:455-465,
//                    clcontext.cpp:877-895)
// Path state stays in the reference's GPUTaskState SoA (geom.h:199-236) so the stages can be
// driven one by one through the C ABI exactly like the reference's enqueueWf*Kernel calls.
#pragma once

#include "flx_bsdf.cuh"
#include "flx_trace.cuh"

#define FLX_BLOCK 256       // streaming stages
#define FLX_TRACE_BLOCK 128 // traversal stages
#define FLX_LOGIC_TILE FLX_BLOCK

// Everything a stage needs besides the scene; passed by value (lives in the constant bank).
struct Frame
{
    Tasks tasks;
    flx_QueueCounters *counters;
    uint32_t *queues[8]; // order of flx_QueueCounters: raygen, extension, shadow, diffuse, glossy, ggxRefl, ggxRefr, delta
    float *pixels;       // tilePixels x float4
    uint8_t *dirty;      // tilePixels: 1 = the accumulator of this pixel changed since the display pass last looked at it
    float *denoiserAlbedo, *denoiserNormal;
    int denoiser;        // accumulate the denoiser feature buffers (the reference's -DUSE_OPTIX_DENOISER, kernel_impl.hpp:53): flx_set_denoiser
    uint32_t *currPixelIdx;
    uint32_t numTasks;
    // image / tile geometry (flx_set_tile): local pixel -> full-image pixel
    uint32_t tilePixels; // pixels owned by this context (= width*height when untiled)
    uint32_t part, nParts, stripeRows;
    float tanHalfFov;    // tan(toRad(0.5f * fov)), evaluated once on the host with flx_tanf (wf_raygen.cl:50)
    int pinholeCamera;   // 1: the thin-lens offset is exactly zero for every sample (aperture 0), decided on the host -- see lens_offset
};

// Thin lens (wf_raygen.cl:59-63, mk_raygen.cl:49-53; disk sample utils.cl:75-80): origin += (worldRadius * aperture) * (right * rx + up * ry)
// with (rx, ry) on the unit disk from the two random numbers the caller has already drawn.  With aperture 0 -- the reference's
// default and every benchmark configuration -- the offset is +-0 for every sample and origin + (+-0) == origin bit for bit,
// PROVIDED no component of the origin is -0.0f (-0 + +0 = +0): the host checks exactly that (flx_update_params) and sets
// pinholeCamera, which skips the pinned double-precision cos / sin here.  They are a sixth of the instructions the fused logic
// kernel issues, executed by the ~5 lanes of a warp that regenerate a path (profiles/r2_base_logic_lines.txt).
FLX_DEV V3 lens_offset(const Frame &fr, const flx_RenderParams &prm, V3 camRight, V3 camUp, float sqrt_r, float th)
{
    if (fr.pinholeCamera)
        return v3(0.0f);
    const float rx = sqrt_r * flx_cosf(th), ry = sqrt_r * flx_sinf(th);
    return (prm.worldRadius * prm.camera.apertureSize) * (camRight * rx + camUp * ry);
}

enum { Q_RAYGEN = 0, Q_EXT, Q_SHADOW, Q_DIFFUSE, Q_GLOSSY, Q_GGXREFL, Q_GGXREFR, Q_DELTA };

FLX_DEV uint32_t *counter_ptr(flx_QueueCounters *c, int q) { return reinterpret_cast<uint32_t *>(c) + q; }

FLX_DEV void write_empty_hit(const Tasks &t, uint32_t g) // EMPTY_HIT(FLT_MAX), geom.h:144 + utils.cl:202-211
{
    t.setv(FLX_S_P, g, v3(0.0f));
    t.setv(FLX_S_N, g, v3(0.0f));
    t.setf(FLX_S_UV, g, 0.0f);
    t.setf(FLX_S_UV + 1, g, 0.0f);
    t.setf(FLX_S_HIT_T, g, 3.402823466e+38f);
    t.setu(FLX_S_HIT_I, g, (uint32_t)-1);
    t.setu(FLX_S_AREA_LIGHT_HIT, g, 0u);
    t.setu(FLX_S_MAT_ID, g, (uint32_t)-1);
}

FLX_DEV void reset_path_fields(const Tasks &t, uint32_t g, float worldRadius) // shared by reset and raygen
{
    t.setv(FLX_S_EI, g, v3(0.0f));
    t.setv(FLX_S_T, g, v3(1.0f));
    t.setu(FLX_S_PATH_LEN, g, 0u);
    t.setu(FLX_S_LAST_SPECULAR, g, 1u);
    t.setf(FLX_S_LAST_PDF_W, g, 1.0f);
    t.setf(FLX_S_LAST_PDF_DIRECT, g, 0.0f);
    t.setf(FLX_S_LAST_PDF_IMPLICIT, g, 0.0f);
    t.setf(FLX_S_LAST_COS_TH, g, 0.0f);
    t.setf(FLX_S_LAST_LIGHT_PICK, g, 1.0f);
    t.setf(FLX_S_SHADOW_RAY_LEN, g, 2.0f * worldRadius);
    t.setu(FLX_S_BACKFACE, g, 0u);
    t.setu(FLX_S_SHADOW_BLOCKED, g, 1u);
    t.setu(FLX_S_FIRST_DIFFUSE, g, 0u);
    t.setv(FLX_S_LAST_EMISSION, g, v3(0.0f));
    t.setv(FLX_S_LAST_BSDF, g, v3(0.0f));
    write_empty_hit(t, g);
}

// ------------------------------------------------------------------------------------------------ reset
__global__ void __launch_bounds__(FLX_BLOCK) k_reset(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm)
{
    const uint32_t gid = blockIdx.x * FLX_BLOCK + threadIdx.x;
    if (gid < fr.tilePixels)
    {
        reinterpret_cast<float4 *>(fr.pixels)[gid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        reinterpret_cast<float4 *>(fr.denoiserNormal)[gid] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        reinterpret_cast<float4 *>(fr.denoiserAlbedo)[gid] = make_float4(0.1f, 0.1f, 0.1f, 0.0f);
        fr.dirty[gid] = 1;
    }
    if (gid >= fr.numTasks)
        return;
    reset_path_fields(fr.tasks, gid, prm.worldRadius);
    fr.tasks.setu(FLX_S_PIXEL_INDEX, gid, 0u);
    // seed = gid (wf_reset.cl:59).  Tiled over several GPUs every part runs the same gid range; with equal seeds all parts would
    // draw the same random sequences for different stripes of the image -- unbiased, but the noise pattern would repeat from
    // stripe to stripe.  So part p starts its seeds at p * numTasks (the hash inside flx_rand decorrelates neighbouring seeds);
    // part 0, and therefore the untiled single-GPU case, is the reference's.
    fr.tasks.setu(FLX_S_SEED, gid, gid + fr.part * fr.numTasks);
    fr.queues[Q_RAYGEN][gid] = gid;
    if (gid == 0)
        fr.counters->raygenQueue = fr.numTasks;
}

// ------------------------------------------------------------------------------------------------ raygen
// local pixel index -> (x, y) in the FULL image: rows are dealt to the parts in stripes of
// stripeRows rows, stripe s -> part s % nParts. Untiled: identity.
FLX_DEV void local_pixel_to_xy(const Frame &fr, uint32_t width, uint32_t local, uint32_t &x, uint32_t &y)
{
    x = local % width;
    const uint32_t ly = local / width;
    const uint32_t stripe = ly / fr.stripeRows, within = ly % fr.stripeRows;
    y = (stripe * fr.nParts + fr.part) * fr.stripeRows + within;
}

// one regenerated camera path (wf_raygen.cl:25-96): pixel, jitter, thin lens, reset of the per-path state
FLX_DEV void raygen_path(const Frame &fr, const flx_RenderParams &prm, uint32_t gid, uint32_t pixelIdx, uint32_t seed)
{
    const Tasks &t = fr.tasks;
    t.setu(FLX_S_PIXEL_INDEX, gid, pixelIdx);
    uint32_t px, py;
    local_pixel_to_xy(fr, prm.width, pixelIdx, px, py);
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

    // thin lens (wf_raygen.cl:59-63; disk sample utils.cl:75-80)
    const V3 fp = camPos + rayDir * prm.camera.focalDist;
    const float sqrt_r = sqrtf(flx_rand(seed));
    const float th = FLX_2PI_F * flx_rand(seed);
    rayOrig = rayOrig + lens_offset(fr, prm, camRight, camUp, sqrt_r, th);
    rayDir = norm3(fp - rayOrig);

    t.setv(FLX_S_ORIG, gid, rayOrig);
    t.setv(FLX_S_DIR, gid, rayDir);
    t.setu(FLX_S_SEED, gid, seed);
    reset_path_fields(t, gid, prm.worldRadius);
}

__global__ void __launch_bounds__(FLX_BLOCK) k_raygen(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm)
{
    const uint32_t count = fr.counters->raygenQueue;
    const uint32_t curr = *fr.currPixelIdx;
    __shared__ uint32_t s_extBase;
    for (uint32_t base = blockIdx.x * FLX_BLOCK; base < count; base += gridDim.x * FLX_BLOCK)
    {
        // every path of this chunk goes to the extension queue: one atomic per CTA reserves the slots up front
        if (threadIdx.x == 0)
            s_extBase = atomicAdd(counter_ptr(fr.counters, Q_EXT), min((uint32_t)FLX_BLOCK, count - base));
        const uint32_t gd = base + threadIdx.x;
        const bool valid = gd < count;
        uint32_t gid = 0;
        if (valid)
        {
        gid = fr.queues[Q_RAYGEN][gd];
        raygen_path(fr, prm, gid, (curr + gd) % fr.tilePixels, fr.tasks.u(FLX_S_SEED, gid));
        }
        __syncthreads();
        if (valid)
            fr.queues[Q_EXT][s_extBase + threadIdx.x] = gid;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ extension rays
struct LocalStack
{
    int s[FLX_STACK_DEPTH];
    FLX_DEV int &operator[](int i) { return s[i]; }
};

// warp-reduce the per-ray work counters and add them to the context totals (instrumented launches only)
FLX_DEV void flush_counts(const RayCount &c, unsigned long long *totals, unsigned rays)
{
    unsigned v[5] = {c.V, c.B, c.T, c.U, rays};
    const unsigned active = __activemask();
#pragma unroll
    for (int k = 0; k < 5; k++)
    {
        unsigned x = v[k];
        x = __reduce_add_sync(active, x);
        if ((threadIdx.x & 31) == (__ffs(active) - 1))
            atomicAdd(totals + k, (unsigned long long)x);
    }
}
FLX_DEV void flush_counts(const NoCount &, unsigned long long *, unsigned) {}

template <class COUNT>
__global__ void __launch_bounds__(FLX_TRACE_BLOCK) k_extrays(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const BvhView bvh,
                                                             const flx_Triangle *tris160, unsigned long long *countTotals)
{
    const uint32_t count = fr.counters->extensionQueue;
    const uint32_t gd = blockIdx.x * FLX_TRACE_BLOCK + threadIdx.x;
    if (gd >= count)
        return;
    const uint32_t gid = fr.queues[Q_EXT][gd];
    const Tasks &t = fr.tasks;
    const V3 o = t.v(FLX_S_ORIG, gid), d = t.v(FLX_S_DIR, gid);

    float tbest = 3.402823466e+38f, ub = 0.0f, vb = 0.0f;
    int tri = -1;
    LocalStack stack;
    COUNT cnt;
    trace_closest(bvh, o, d, tbest, ub, vb, tri, stack, cnt);
    flush_counts(cnt, countTotals, 1u);

    // hit record (bvh.cl:271-279): attributes of the winning triangle, fetched once
    V3 P = v3(0.0f), N = v3(0.0f);
    float tu = 0.0f, tv = 0.0f;
    int matId = -1, lightHit = 0;
    if (tri >= 0)
    {
        P = o + tbest * d;
        hit_attributes(bvh, tri, ub, vb, N, tu, tv, matId);
    }
    if (prm.sampleImpl && prm.useAreaLight) // wf_extrays.cl:29, intersect.cl:124-155
    {
        if (light_quad(prm.areaLight, o, d, tbest))
        {
            lightHit = 1;
            P = o + tbest * d;
            N = v3(prm.areaLight.N);
            tri = 0;
            matId = 0;
        }
    }
    t.setu(FLX_S_PATH_LEN, gid, t.u(FLX_S_PATH_LEN, gid) + 1u);
    t.setv(FLX_S_P, gid, P);
    t.setv(FLX_S_N, gid, N);
    t.setf(FLX_S_UV, gid, tu);
    t.setf(FLX_S_UV + 1, gid, tv);
    t.setf(FLX_S_HIT_T, gid, tbest);
    t.setu(FLX_S_HIT_I, gid, (uint32_t)tri);
    t.setu(FLX_S_AREA_LIGHT_HIT, gid, (uint32_t)lightHit);
    t.setu(FLX_S_MAT_ID, gid, (uint32_t)matId);
}

// ------------------------------------------------------------------------------------------------ shadow rays
template <class COUNT>
__global__ void __launch_bounds__(FLX_TRACE_BLOCK) k_shadowrays(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const BvhView bvh,
                                                                unsigned long long *countTotals)
{
    const uint32_t count = fr.counters->shadowQueue;
    const uint32_t gd = blockIdx.x * FLX_TRACE_BLOCK + threadIdx.x;
    if (gd >= count)
        return;
    const uint32_t gid = fr.queues[Q_SHADOW][gd];
    const Tasks &t = fr.tasks;
    const V3 o = t.v(FLX_S_SHADOW_ORIG, gid), d = t.v(FLX_S_SHADOW_DIR, gid);
    const float lenL = t.f(FLX_S_SHADOW_RAY_LEN, gid);
    bool occluded = false;
    if (prm.useAreaLight) // the light quad is tested first and blocks (wf_shadowrays.cl:29-31)
    {
        float tl = lenL;
        occluded = light_quad(prm.areaLight, o, d, tl);
    }
    COUNT cnt;
    if (!occluded)
    {
        LocalStack stack;
        occluded = trace_any(bvh, o, d, lenL, stack, cnt);
    }
    flush_counts(cnt, countTotals, 1u);
    t.setu(FLX_S_SHADOW_BLOCKED, gid, occluded ? 1u : 0u);
}

// ------------------------------------------------------------------------------------------------ logic
// Stable compaction state for the raygen queue: one 64-bit word per 256-path tile,
// (status << 32 | value); status 0 = not ready, 1 = tile aggregate, 2 = inclusive prefix.
struct ScanState
{
    unsigned long long *tiles;
    uint32_t *ticket;
};
#define SCAN_AGG (1ull << 32)
#define SCAN_PREFIX (2ull << 32)

FLX_DEV unsigned long long ld_relaxed(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
FLX_DEV void st_relaxed(unsigned long long *p, unsigned long long v) { asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }

FLX_DEV float luminance3(V3 v) { return (0.212671f * v.x + 0.715160f * v.y) + 0.072169f * v.z; } // utils.cl:237-240

// one path through its material (wf_mat_*.cl:33-62): BSDF value and pdf toward the pending light sample, then the continuation ray
template <int MASK>
FLX_DEV void material_path(const Tasks &t, const SceneView &sc, uint32_t gid, const Surface &s, const Mat &mat, bool backface, V3 dirIn, V3 L, V3 oldT, uint32_t seed)
{
    // BSDF value and pdf toward the pending light sample (wf_mat_*.cl:33-36)
    const V3 bsdfNEE = bxdf_eval<MASK>(s, mat, backface, sc, dirIn, L);
    const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf<MASK>(s, mat, backface, sc, dirIn, L));
    t.setv(FLX_S_LAST_BSDF, gid, bsdfNEE);
    t.setf(FLX_S_LAST_PDF_IMPLICIT, gid, bsdfPdfW);

    // continuation (wf_mat_*.cl:38-62)
    float pdfW = 0.0f;
    V3 newDir = v3(0.0f);
    const V3 bsdf = bxdf_sample<MASK>(s, mat, backface, sc, dirIn, newDir, pdfW, seed);
    const float costh = dot3(s.N, norm3(newDir));
    V3 newT = v3(0.0f);
    if (!(pdfW == 0.0f || is_zero3(bsdf)))
        newT = ((oldT * bsdf) * costh) / pdfW;
    const V3 orig = s.P + 1e-4f * newDir;
    t.setv(FLX_S_LAST_T, gid, oldT);
    t.setv(FLX_S_T, gid, newT);
    t.setv(FLX_S_ORIG, gid, orig);
    t.setv(FLX_S_DIR, gid, newDir);
    t.setf(FLX_S_LAST_PDF_W, gid, pdfW);
    t.setu(FLX_S_SEED, gid, seed);
    t.setu(FLX_S_LAST_SPECULAR, gid, (mat.type & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC)) != 0 ? 1u : 0u);
}


// Denoiser feature buffers (reference: the USE_OPTIX_DENOISER blocks, src/wf_logic.cl:186-209, src/mk_next_vertex.cl:60-70,
// src/mk_sample_bsdf.cl:56-66).  First-hit shading normal in camera space: rows right, up, -dir of the camera frame (a rotation,
// so its inverse transpose is itself), w counts the samples; albedo = Kd (texture or constant, NOT gamma-corrected) at the first
// non-singular vertex of the path, once per path (firstDiffuseHit).
FLX_DEV float4 denoiser_normal(const flx_RenderParams &prm, V3 N)
{
    const V3 r1 = v3(prm.camera.right), r2 = v3(prm.camera.up), r3 = -v3(prm.camera.dir);
    return make_float4(dot3(r1, N), dot3(r2, N), dot3(r3, N), 1.0f);
}
// wavefront flavour: several paths in flight may hold the same pixel, so the adds are atomic (the reference builds with
// -DFLT_FLOAT_ATOMICS, clcontext.cpp:145).  Out of line: it is off in the default configuration and must not cost registers there.
__device__ __noinline__ void denoiser_aov_wf(const Frame &fr, const flx_RenderParams &prm, const SceneView &sc, uint32_t gid, uint32_t len, uint32_t pixIdx, V3 N, const Mat &mat,
                                            float u, float v)
{
    if (len == 1u)
        atomicAdd(reinterpret_cast<float4 *>(fr.denoiserNormal) + pixIdx, denoiser_normal(prm, N));
    const bool isDiffuse = (mat.type & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC)) == 0;
    if (isDiffuse && fr.tasks.u(FLX_S_FIRST_DIFFUSE, gid) == 0u)
    {
        fr.tasks.setu(FLX_S_FIRST_DIFFUSE, gid, 1u);
        const V3 albedo = mat_float3(mat.Kd, u, v, mat.map_Kd, sc);
        atomicAdd(reinterpret_cast<float4 *>(fr.denoiserAlbedo) + pixIdx, make_float4(albedo.x, albedo.y, albedo.z, 1.0f));
    }
}

// Fused stages: the same thread also does what wf_raygen (for a path it has just terminated) and wf_mat_* (for a path it sends on)
// would do next, instead of leaving that to two more passes over the path state.  Nothing crosses paths between those three
// stages except the queue counters and the raygen-queue rank (known here from the look-back scan), so the state, the queues and
// the counters after this one launch are exactly those after logic + raygen + materials.  Used by flx_render (and by the
// per-stage ABI when the three calls arrive back to back, flx_api.cu); what it saves is the sparse second and third pass:
// raygen touches ~1 in 5 paths (one 32-byte sector per 4 bytes used), the material stage re-reads what logic had in registers.
// FUSE: 0 = wf_logic only; 1 = + the camera-ray part; 2 = + the material part as well.  The material part is worth fusing when all
// materials go through one kernel anyway (the reference's single-queue mode, wavefrontAllMaterials): with per-type queues the
// point of the separate kernels is that a warp sees ONE BSDF, and folding them into this kernel puts up to five heavy lobes into
// every warp (Country Kitchen: 0.61 ms for the three kernels, 3.8 ms fully fused) -- so with wfSeparateQueues only level 1 is used.
// LT: paths per tile = threads per CTA (256, or 128: half as many warps to wait for at each of the four barriers)
// MATMASK: the BSDF lobes the fused material part compiles in.  The reference builds its all-materials kernel with only the lobes the
// scene's materials use ("Only handle material types that exist in scene", src/kernel_impl.hpp:261-266, getBxdfDefines in
// src/utils.cpp:93-113); here the host picks the instantiation from the uploaded materials' types: all lobes, or diffuse only (Conference).
#define FLX_ALL_BXDF (FLX_BXDF_DIFFUSE | FLX_BXDF_GLOSSY | FLX_BXDF_GGX_ROUGH_REFLECTION | FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_GGX_ROUGH_DIELECTRIC | FLX_BXDF_IDEAL_DIELECTRIC | FLX_BXDF_EMISSIVE)
#define FLX_CHEAP_BXDF (FLX_BXDF_DIFFUSE | FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC) // lobes without GGX / glossy machinery
template <bool SEPARATE_QUEUES, int MIN_BLOCKS, int FUSE = 0, int LT = FLX_LOGIC_TILE, int MATMASK = FLX_ALL_BXDF>
__global__ void __launch_bounds__(LT, MIN_BLOCKS) k_logic(const __grid_constant__ Frame fr, const __grid_constant__ flx_RenderParams prm, const SceneView sc,
                                                     const ScanState scan, const uint32_t maxId)
{
    __shared__ uint32_t s_tile;
    __shared__ uint32_t s_warpCount[LT / 32];
    __shared__ uint32_t s_base;
    if (threadIdx.x == 0)
        s_tile = atomicAdd(scan.ticket, 1u); // tiles are numbered in the order blocks start, so look-back never waits on an unscheduled block
    __syncthreads();
    const uint32_t tile = s_tile;
    const uint32_t gid = tile * LT + threadIdx.x;
    const bool live = gid < maxId;
    const Tasks &t = fr.tasks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

    // ---- all loads up front: one DRAM round trip instead of a chain of dependent ones (this stage is latency-bound:
    //      ncu showed 17 warps stalled on the long scoreboard per issued instruction at 36 % occupancy).  Fields that only
    //      some paths need (the pending light sample, the hit record) are fetched for all: ~10 % more bytes, 3 fewer trips.
    uint32_t seed = 0, len = 0, pixIdx = 0;
    V3 T = v3(0.0f), rayOrig = v3(0.0f), rayDir = v3(0.0f), Ei = v3(0.0f);
    V3 hP = v3(0.0f), hN = v3(0.0f), neeEmission = v3(0.0f), neeBsdf = v3(0.0f), neeT = v3(0.0f);
    float hU = 0.0f, hV = 0.0f, lastPdfW = 0.0f, lastLightPick = 0.0f, neeCosTh = 0.0f, neePdfDirect = 0.0f, neePdfImplicit = 0.0f;
    int hitI = -1, hitLight = 0, matId = 0;
    bool lastSpecular = false, blocked = true;
    bool terminate = false;
    if (live)
    {
        seed = t.u(FLX_S_SEED, gid);
        len = t.u(FLX_S_PATH_LEN, gid);
        hitI = (int)t.u(FLX_S_HIT_I, gid);
        hitLight = (int)t.u(FLX_S_AREA_LIGHT_HIT, gid);
        matId = (int)t.u(FLX_S_MAT_ID, gid);
        rayOrig = t.v(FLX_S_ORIG, gid);
        rayDir = t.v(FLX_S_DIR, gid);
        T = t.v(FLX_S_T, gid);
        Ei = t.v(FLX_S_EI, gid);
        lastPdfW = t.f(FLX_S_LAST_PDF_W, gid);
        lastSpecular = t.u(FLX_S_LAST_SPECULAR, gid) != 0u;
        lastLightPick = t.f(FLX_S_LAST_LIGHT_PICK, gid);
        blocked = t.u(FLX_S_SHADOW_BLOCKED, gid) != 0u;
        pixIdx = t.u(FLX_S_PIXEL_INDEX, gid);
        hP = t.v(FLX_S_P, gid);
        hN = t.v(FLX_S_N, gid);
        hU = t.f(FLX_S_UV, gid);
        hV = t.f(FLX_S_UV + 1, gid);
        neeEmission = t.v(FLX_S_LAST_EMISSION, gid);
        neeBsdf = t.v(FLX_S_LAST_BSDF, gid);
        neeT = t.v(FLX_S_LAST_T, gid);
        neeCosTh = t.f(FLX_S_LAST_COS_TH, gid);
        neePdfDirect = t.f(FLX_S_LAST_PDF_DIRECT, gid);
        neePdfImplicit = t.f(FLX_S_LAST_PDF_IMPLICIT, gid);
    }

    // ---- phase 1: decide termination (wf_logic.cl:48-127) -------------------------------------------------
    if (live)
    {
        terminate = (len >= prm.maxBounces + 1u);
        if (terminate && prm.useRoulette)
        {
            const float contProb = fminf(fmaxf(luminance3(T), 0.01f), 0.5f);
            terminate = (flx_rand(seed) > contProb);
            T = T / contProb;
            t.setv(FLX_S_T, gid, T);
        }
        if (is_zero3(T) || lastPdfW == 0.0f)
            terminate = true;

        if (hitI < 0 && !terminate) // escaped: implicit environment sample
        {
            float weight = 1.0f;
            V3 bg = v3(0.0f);
            if (prm.useEnvMap && (len == 1u || prm.sampleImpl))
                bg = env_eval_dir(sc, rayDir) * prm.envMapStrength;
            if (prm.sampleImpl && prm.sampleExpl && prm.useEnvMap && len > 1u && !lastSpecular)
            {
                const float directPdfW = env_pdf(sc, rayDir);
                weight = (lastPdfW * lastLightPick) / (lastPdfW * lastLightPick + directPdfW);
            }
            Ei = Ei + (weight * T) * bg;
            terminate = true;
        }
        else if (hitLight && !terminate) // implicit area-light sample
        {
            float misWeight = 1.0f;
            if (prm.sampleExpl && len > 1u && !lastSpecular)
            {
                const float directPdfA = 1.0f / (4.0f * prm.areaLight.size.x * prm.areaLight.size.y);
                const float dist = len3(hP - rayOrig);
                const float cosine = dot3(norm3(-rayDir), hN);
                const float directPdfW = directPdfA * (dist * dist) / fabsf(cosine); // pdfAtoW, utils.cl:197-200
                misWeight = lastPdfW / (lastPdfW + directPdfW * lastLightPick);
            }
            Ei = Ei + (T * misWeight) * v3(prm.areaLight.E);
            terminate = true;
        }
    }

    // publish this tile's count of terminated paths as early as possible
    const unsigned termMask = __ballot_sync(0xffffffffu, live && terminate);
    if (lane == 0)
        s_warpCount[warp] = __popc(termMask);
    __syncthreads();
    uint32_t tileCount = 0, warpBase = 0;
#pragma unroll
    for (int w = 0; w < LT / 32; w++)
    {
        if (w == warp)
            warpBase = tileCount;
        tileCount += s_warpCount[w];
    }
    if (threadIdx.x == 0)
        st_relaxed(scan.tiles + tile, (tile == 0 ? SCAN_PREFIX : SCAN_AGG) | tileCount);

    // ---- phase 2: previous vertex's light sample, splat, next-event estimation (wf_logic.cl:129-303) -------
    if (live)
    {
        bool eiDirty = terminate && (hitI < 0 || hitLight); // Ei was updated above only on the implicit-hit branches
        // (a terminated path whose Ei did not change keeps its stored value; writing the same bits back is harmless, so
        //  the flag only saves stores)
        if (!blocked)
        {
            float weight = 1.0f;
            if (prm.sampleImpl)
                weight = (neePdfDirect * lastLightPick) / (neePdfDirect * lastLightPick + neePdfImplicit);
            const V3 contrib = ((((neeBsdf * neeT) * neeEmission) * weight) * neeCosTh) / (lastLightPick * neePdfDirect);
            Ei = Ei + contrib;
            eiDirty = true;
        }
        if (eiDirty)
            t.setv(FLX_S_EI, gid, Ei);
        if (terminate)
        {
            if (len > 0u)
            {
                atomicAdd(reinterpret_cast<float4 *>(fr.pixels) + pixIdx, make_float4(Ei.x, Ei.y, Ei.z, 1.0f)); // one 128-bit reduction
                fr.dirty[pixIdx] = 1;
            }
            if (FUSE == 0) // fused: the camera-ray part below carries the seed on and stores it
                t.setu(FLX_S_SEED, gid, seed);
        }
    }

    bool toMaterial = false, pushShadow = false;
    int matType = 0;
    // FUSED: what the material part at the end of the kernel needs (dead code otherwise)
    Mat fMat;
    Surface fS;
    bool fBackface = false, fHaveL = false;
    V3 fL = v3(0.0f);
    if (live && !terminate)
    {
        Surface s;
        s.P = hP;
        s.N = hN;
        s.u = hU;
        s.v = hV;
        s.tri = hitI;
        const Mat mat = load_material(sc.materials, matId, sc.kdGamma);
        V3 N = shading_normal(s, mat, sc);
        const bool backface = dot3(N, rayDir) > 0.0f;
        if (backface)
            N = N * -1.0f;
        const V3 orig = s.P - 1e-3f * rayDir;
        if (fr.denoiser) // wf_logic.cl:186-209 (before the hit record is written back, like there)
            denoiser_aov_wf(fr, prm, sc, gid, len, pixIdx, N, mat, hU, hV);
        t.setv(FLX_S_N, gid, N);
        t.setu(FLX_S_BACKFACE, gid, backface ? 1u : 0u);

        bool wroteL = false; // FUSED: the light direction the material stage would read back from the shadowDir slot
        V3 newL = v3(0.0f);
        const bool singular = (mat.type & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC)) != 0;
        if (prm.sampleExpl && !singular)
        {
            const uint32_t nLights = prm.useEnvMap + prm.useAreaLight;
            const float envMapProb = (float)prm.useEnvMap / (float)(nLights > 1u ? nLights : 1u);
            const bool useEnv = flx_rand(seed) < envMapProb;
            const bool useArea = !useEnv && prm.useAreaLight;
            if (useEnv && prm.useEnvMap)
            {
                V3 L;
                float directPdfW = 0.0f;
                env_sample_alias(sc, flx_rand(seed), L, directPdfW);
                const float lenL = 2.0f * prm.worldRadius;
                L = norm3(L);
                const float cosTh = fmaxf(0.0f, dot3(L, N));
                const V3 Li = env_eval_dir(sc, L) * prm.envMapStrength;
                t.setv(FLX_S_SHADOW_ORIG, gid, orig);
                t.setv(FLX_S_SHADOW_DIR, gid, L);
                t.setf(FLX_S_SHADOW_RAY_LEN, gid, lenL);
                t.setf(FLX_S_LAST_PDF_DIRECT, gid, directPdfW);
                t.setf(FLX_S_LAST_COS_TH, gid, cosTh);
                t.setf(FLX_S_LAST_LIGHT_PICK, gid, envMapProb);
                t.setv(FLX_S_LAST_EMISSION, gid, Li);
                pushShadow = true;
                wroteL = true;
                newL = L;
            }
            if (useArea)
            {
                const float lightPickProb = 1.0f - envMapProb;
                const flx_AreaLight &A = prm.areaLight;
                const float directPdfA = 1.0f / (4.0f * A.size.x * A.size.y); // sampleAreaLight, utils.cl:226-234
                V3 posL = v3(A.pos);
                const float r1 = 2.0f * flx_rand(seed) - 1.0f;
                const float r2 = 2.0f * flx_rand(seed) - 1.0f;
                posL = posL + (r1 * A.size.x) * v3(A.right);
                posL = posL + (r2 * A.size.y) * v3(A.up);
                V3 L = posL - orig;
                const float lenL = len3(L) * 0.995f;
                L = norm3(L);
                const float cosLight = fmaxf(dot3(v3(A.N), -L), 0.0f);
                if (cosLight > 0.0f)
                {
                    const float directPdfW = directPdfA * (lenL * lenL) / fabsf(cosLight);
                    const float cosTh = fmaxf(0.0f, dot3(L, N));
                    t.setv(FLX_S_SHADOW_ORIG, gid, orig);
                    t.setv(FLX_S_SHADOW_DIR, gid, L);
                    t.setf(FLX_S_SHADOW_RAY_LEN, gid, lenL);
                    t.setf(FLX_S_LAST_PDF_DIRECT, gid, directPdfW);
                    t.setf(FLX_S_LAST_COS_TH, gid, cosTh);
                    t.setf(FLX_S_LAST_LIGHT_PICK, gid, lightPickProb);
                    t.setv(FLX_S_LAST_EMISSION, gid, v3(A.E));
                    pushShadow = true;
                    wroteL = true;
                    newL = L;
                }
                else
                    t.setu(FLX_S_SHADOW_BLOCKED, gid, 1u);
            }
        }
        toMaterial = true;
        matType = mat.type;
        if (FUSE == 2) // the material part runs at the very end, after the last barrier (threads with heavy BSDFs would otherwise hold up their CTA)
        {
            fMat = mat;
            fS = s;
            fS.N = N; // the material stage reads the flipped shading normal logic has just stored
            fBackface = backface;
            fHaveL = wroteL;
            fL = newL;
        }
        else
            t.setu(FLX_S_SEED, gid, seed);
    }

    // ---- shadow + material queues: ONE atomic per queue per 256-path tile.  All paths of a tile hammer the same
    //      counter word, and same-address atomics serialise in the L2 slice that owns it; aggregating over the CTA
    //      (ballot -> per-warp counts in shared memory -> one atomicAdd by one thread per queue) cuts them 8x versus the
    //      per-warp aggregation the reference's NVIDIA path does (wf_logic.cl:459-519, ptx_asm.cl:83-111).
    {
        constexpr int NQ = SEPARATE_QUEUES ? 6 : 2; // slot 0: shadow, slots 1..: material queues; FUSED: slot NQ = extension queue
        constexpr int NW = LT / 32;
        __shared__ uint32_t s_cnt[7][NW];
        __shared__ uint32_t s_qbase[7];
        int q = -1; // material queue slot of this path (1-based), -1: none
        if (toMaterial)
        {
            if (!SEPARATE_QUEUES) q = 1;
            else if (matType == FLX_BXDF_DIFFUSE) q = 1;
            else if (matType == FLX_BXDF_GLOSSY) q = 2;
            else if (matType == FLX_BXDF_GGX_ROUGH_REFLECTION) q = 3;
            else if (matType == FLX_BXDF_GGX_ROUGH_DIELECTRIC) q = 4;
            else if (matType == FLX_BXDF_IDEAL_REFLECTION || matType == FLX_BXDF_IDEAL_DIELECTRIC) q = 5;
            // any other type is dropped, as in the reference (wf_logic.cl:362-364)
        }
        unsigned myMask = 0u, shadowMask = __ballot_sync(0xffffffffu, pushShadow);
        if (lane == 0)
            s_cnt[0][warp] = __popc(shadowMask);
#pragma unroll
        for (int k = 1; k < NQ; k++)
        {
            const unsigned m = __ballot_sync(0xffffffffu, q == k);
            if (lane == 0)
                s_cnt[k][warp] = __popc(m);
            if (q == k)
                myMask = m;
        }
        // FUSED: every path that was regenerated or sent through its material is an extension ray of this iteration
        // (fused material part with per-type queues: only the lobes compiled in are done here, the others stay with their queue's kernel)
        const bool pushExt = FUSE >= 1 && live && (terminate || (FUSE == 2 && q > 0 && (!SEPARATE_QUEUES || (matType & MATMASK) != 0)));
        const unsigned extMask = FUSE >= 1 ? __ballot_sync(0xffffffffu, pushExt) : 0u;
        if (FUSE >= 1 && lane == 0)
            s_cnt[NQ][warp] = __popc(extMask);
        __syncthreads();
        if (threadIdx.x < NQ + (FUSE >= 1 ? 1 : 0))
        {
            // one thread per queue: per-warp counts -> per-warp offsets (in place), then ONE atomic for the tile's total
            uint32_t total = 0;
#pragma unroll
            for (int w = 0; w < NW; w++)
            {
                const uint32_t c = s_cnt[threadIdx.x][w];
                s_cnt[threadIdx.x][w] = total;
                total += c;
            }
            const int queueId = threadIdx.x == 0 ? Q_SHADOW : (threadIdx.x == NQ ? Q_EXT : Q_DIFFUSE + (int)threadIdx.x - 1);
            s_qbase[threadIdx.x] = total ? atomicAdd(counter_ptr(fr.counters, queueId), total) : 0u;
        }
        __syncthreads();
        const unsigned below = (1u << lane) - 1u;
        if (pushShadow)
            fr.queues[Q_SHADOW][s_qbase[0] + s_cnt[0][warp] + __popc(shadowMask & below)] = gid;
        if (q > 0)
            fr.queues[Q_DIFFUSE + q - 1][s_qbase[q] + s_cnt[q][warp] + __popc(myMask & below)] = gid;
        if (pushExt)
            fr.queues[Q_EXT][s_qbase[NQ] + s_cnt[NQ][warp] + __popc(extMask & below)] = gid;
    }

    // ---- raygen queue in ascending path order: decoupled look-back over the tile words ----------------------
    if (warp == 0 && tile > 0)
    {
        uint32_t running = 0;
        int look = (int)tile - 1 - lane;
        while (true)
        {
            unsigned long long st = SCAN_PREFIX; // tiles before tile 0 contribute an empty inclusive prefix
            if (look >= 0)
            {
                st = ld_relaxed(scan.tiles + look);
                while ((st >> 32) == 0ull)
                    st = ld_relaxed(scan.tiles + look);
            }
            const unsigned isPrefix = __ballot_sync(0xffffffffu, (st >> 32) == 2ull);
            const int firstP = isPrefix ? (__ffs(isPrefix) - 1) : 31;
            uint32_t v = (lane <= firstP) ? (uint32_t)(st & 0xffffffffull) : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                v += __shfl_xor_sync(0xffffffffu, v, o);
            running += v;
            if (isPrefix)
                break;
            look -= 32;
        }
        if (lane == 0)
        {
            s_base = running;
            st_relaxed(scan.tiles + tile, SCAN_PREFIX | (unsigned long long)(running + tileCount));
        }
    }
    else if (threadIdx.x == 0 && tile == 0)
        s_base = 0;
    __syncthreads();
    if (live && terminate)
    {
        const uint32_t rank = s_base + warpBase + __popc(termMask & ((1u << lane) - 1u));
        fr.queues[Q_RAYGEN][rank] = gid;
        if (FUSE >= 1) // wf_raygen for this path: its queue position is the rank just computed (wf_raygen.cl:25)
            raygen_path(fr, prm, gid, (*fr.currPixelIdx + rank) % fr.tilePixels, seed);
    }
    if (FUSE == 2 && toMaterial)
    {
        bool hasQueue = true;
        if (SEPARATE_QUEUES) // a type without a queue is dropped, as in the reference (wf_logic.cl:362-364)
            hasQueue = matType == FLX_BXDF_DIFFUSE || matType == FLX_BXDF_GLOSSY || matType == FLX_BXDF_GGX_ROUGH_REFLECTION || matType == FLX_BXDF_GGX_ROUGH_DIELECTRIC ||
                       matType == FLX_BXDF_IDEAL_REFLECTION || matType == FLX_BXDF_IDEAL_DIELECTRIC;
        if (hasQueue && (!SEPARATE_QUEUES || (matType & MATMASK) != 0))
        {
            const V3 L = fHaveL ? fL : t.v(FLX_S_SHADOW_DIR, gid); // no new light sample: whatever an earlier vertex left in the slot
            material_path<MATMASK>(t, sc, gid, fS, fMat, fBackface, rayDir, L, T, seed);
        }
        else
            t.setu(FLX_S_SEED, gid, seed);
    }
    // the last tile knows the total
    if (threadIdx.x == 0 && (tile + 1u) * LT >= maxId && tile * LT < maxId)
        atomicAdd(counter_ptr(fr.counters, Q_RAYGEN), s_base + tileCount);
}

// ------------------------------------------------------------------------------------------------ materials
template <int MASK>
__global__ void __launch_bounds__(FLX_BLOCK) k_material(const __grid_constant__ Frame fr, const SceneView sc, const int queue)
{
    const uint32_t count = *counter_ptr(fr.counters, queue);
    const Tasks &t = fr.tasks;
    __shared__ uint32_t s_extBase;
    for (uint32_t base = blockIdx.x * FLX_BLOCK; base < count; base += gridDim.x * FLX_BLOCK)
    {
        if (threadIdx.x == 0) // one atomic per CTA reserves this chunk's extension-queue slots
            s_extBase = atomicAdd(counter_ptr(fr.counters, Q_EXT), min((uint32_t)FLX_BLOCK, count - base));
        const uint32_t gd = base + threadIdx.x;
        const bool valid = gd < count;
        uint32_t gid = 0;
        if (valid)
        {
        gid = fr.queues[queue][gd];
        uint32_t seed = t.u(FLX_S_SEED, gid);
        Surface s;
        s.P = t.v(FLX_S_P, gid);
        s.N = t.v(FLX_S_N, gid);
        s.u = t.f(FLX_S_UV, gid);
        s.v = t.f(FLX_S_UV + 1, gid);
        s.tri = (int)t.u(FLX_S_HIT_I, gid);
        const Mat mat = load_material(sc.materials, (int)t.u(FLX_S_MAT_ID, gid), sc.kdGamma);
        const bool backface = t.u(FLX_S_BACKFACE, gid) != 0u;
        const V3 dirIn = t.v(FLX_S_DIR, gid); // points toward the surface
        const V3 L = t.v(FLX_S_SHADOW_DIR, gid);
        const V3 oldT = t.v(FLX_S_T, gid);

        material_path<MASK>(t, sc, gid, s, mat, backface, dirIn, L, oldT, seed);

        }
        __syncthreads();
        if (valid)
            fr.queues[Q_EXT][s_extBase + threadIdx.x] = gid;
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ post-process
// The display pass the reference runs at the end of every loop iteration (src/mk_postprocess.cl:7-55 with
// src/tonemap.cl:3-26; enqueued at src/tracer.cpp:302 and :447): normalise by the sample count, exposure, Reinhard or
// Uncharted-2 tone map, gamma 1/2.2.  Pure streaming: 16 B in, 16 B out per pixel.
FLX_DEV float uc2_curve(float x)
{
    const float A = 0.22, B = 0.30, C = 0.10, D = 0.20, E = 0.01, F = 0.30; // double literals rounded to float, as in the reference
    return (x * (A * x + C * B) + D * E) / (x * (A * x + B) + D * F) - E / F;
}
// Only pixels whose accumulator changed since the last pass are recomputed (`dirty`, set by every kernel that writes the
// accumulator; `all` after anything else that invalidates the preview: new image, new exposure / operator): an iteration splats
// into at most a fifth of the pixels, and the three pinned double-precision pow per pixel are what this pass costs.  The
// preview buffer after the pass is the same as if every pixel had been recomputed.
// the denoiser feature buffers as the display pass hands them on (mk_postprocess.cl:49-54): divided by their sample count once
// that exceeds 1 (sic: "> 1.0f", so a single sample is passed through as it is -- same value)
__global__ void __launch_bounds__(FLX_BLOCK) k_postprocess_aovs(const float4 *__restrict__ normal, const float4 *__restrict__ albedo, float4 *__restrict__ normalOut,
                                                                float4 *__restrict__ albedoOut, uint32_t nPixels)
{
    for (uint32_t i = blockIdx.x * FLX_BLOCK + threadIdx.x; i < nPixels; i += gridDim.x * FLX_BLOCK)
    {
        const float4 n = normal[i], a = albedo[i];
        normalOut[i] = n.w > 1.0f ? make_float4(n.x / n.w, n.y / n.w, n.z / n.w, n.w / n.w) : n;
        albedoOut[i] = a.w > 1.0f ? make_float4(a.x / a.w, a.y / a.w, a.z / a.w, a.w / a.w) : a;
    }
}

__global__ void __launch_bounds__(FLX_BLOCK) k_postprocess(const float4 *__restrict__ pixels, float4 *__restrict__ preview, uint8_t *__restrict__ dirty, const int all,
                                                           uint32_t nPixels, float exposure, uint32_t tmOperator)
{
    const float W = 11.2, exposureBias = 2.0;
    const float white = uc2_curve(W);
    for (uint32_t i = blockIdx.x * FLX_BLOCK + threadIdx.x; i < nPixels; i += gridDim.x * FLX_BLOCK)
    {
        if (dirty) // (null: the full image gathered from all ranks, which has no change marks)
        {
            if (!all && !dirty[i])
                continue;
            dirty[i] = 0;
        }
        float4 c = pixels[i];
        if (c.w > 0.0f)
        {
            const float w = c.w;
            c = make_float4(c.x / w, c.y / w, c.z / w, c.w / w);
        }
        float r = c.x * exposure, g = c.y * exposure, b = c.z * exposure;
        if (tmOperator == 1u)
        {
            r = r / (1.0f + r);
            g = g / (1.0f + g);
            b = b / (1.0f + b);
        }
        if (tmOperator == 2u)
        {
            r = uc2_curve(exposureBias * r) / white;
            g = uc2_curve(exposureBias * g) / white;
            b = uc2_curve(exposureBias * b) / white;
        }
        const float ig = 1.0f / 2.2f;
        preview[i] = make_float4(flx_powf(r, ig), flx_powf(g, ig), flx_powf(b, ig), c.w);
    }
}

// ------------------------------------------------------------------------------------------------ between iterations
struct IterationState
{
    flx_QueueCounters *counters;
    flx_QueueCounters *snapshot; // counters as they were after the material stage (enqueueGetCounters point); may be `counters` itself
    uint32_t *fetch;             // the two traversal kernels' queue-fetch counters, zeroed here for the next iteration (null: leave them)
    unsigned long long *scanTiles; // the logic kernel's tile status words to zero (nScanTiles of them; 0: leave them) and its ticket
    uint32_t nScanTiles;
    uint32_t *scanTicket;
    flx_RenderStats64 *stats;
    uint32_t *currPixelIdx;
    uint32_t tilePixels;
};
// single thread: stats += snapshot (tracer.cpp:455-462), pixelIdx advance (clcontext.cpp:891-895), clear (877-883)
__global__ void k_end_iteration(const IterationState it)
{
    // the logic kernel's scan state (one status word per tile + the tile ticket) back to zero for the next iteration: these words and the
    // fetch counters below were three memset nodes in front of the next iteration's kernels
    for (uint32_t i = threadIdx.x; i < it.nScanTiles; i += blockDim.x)
        it.scanTiles[i] = 0ull;
    if (threadIdx.x != 0 || blockIdx.x != 0)
        return;
    if (it.scanTicket)
        *it.scanTicket = 0u;
    const flx_QueueCounters c = *it.snapshot;
    it.stats->extensionRays += c.extensionQueue;
    it.stats->shadowRays += c.shadowQueue;
    it.stats->primaryRays += c.raygenQueue;
    it.stats->samples += c.raygenQueue;
    it.stats->iterations += 1;
    *it.currPixelIdx = (uint32_t)(((uint64_t)*it.currPixelIdx + c.raygenQueue) % it.tilePixels);
    flx_QueueCounters z = {0, 0, 0, 0, 0, 0, 0, 0};
    *it.counters = z;
    if (it.fetch)
        it.fetch[0] = it.fetch[1] = 0u;
}
