// ref_kernels.cpp -- TEST INFRASTRUCTURE (oracle).
// One translation unit per reference kernel file: build_ref.py compiles this file once per
// kernel with -DREF_TU_<name> and an include path that holds the "(float3)(" -> "float3("
// transformed copy of /root/reference/src/*.cl (made in a temporary directory at build time;
// no reference source is stored in this repository).  The kernel body is the reference's own.
// Each exported function runs the NDRange [begin, end) one work-item at a time; with
// -DSHIM_PARALLEL the loop is an OpenMP parallel-for (CPU baseline), otherwise it is serial
// and therefore deterministic (parity oracle).
#include "cl_shim.hpp"
#include "ref_abi.h"

namespace
{
#if defined(REF_TU_RESET)
#include "wf_reset.cl"
#elif defined(REF_TU_RAYGEN)
#include "wf_raygen.cl"
#elif defined(REF_TU_EXT)
#include "wf_extrays.cl"
#elif defined(REF_TU_SHADOW)
#include "wf_shadowrays.cl"
#elif defined(REF_TU_LOGIC_SINGLE) || defined(REF_TU_LOGIC_SEPARATE)
#include "wf_logic.cl"
#elif defined(REF_TU_MAT_ALL)
#include "wf_mat_all.cl"
#elif defined(REF_TU_MAT_DIFFUSE)
#include "wf_mat_diffuse.cl"
#elif defined(REF_TU_MAT_GLOSSY)
#include "wf_mat_glossy.cl"
#elif defined(REF_TU_MAT_GGX_REFL)
#include "wf_mat_ggx_reflection.cl"
#elif defined(REF_TU_MAT_GGX_REFR)
#include "wf_mat_ggx_refraction.cl"
#elif defined(REF_TU_MAT_DELTA)
#include "wf_mat_delta.cl"
#elif defined(REF_TU_POSTPROCESS)
#include "mk_postprocess.cl"
#elif defined(REF_TU_MK_RESET)
#include "mk_reset.cl"
#elif defined(REF_TU_MK_RAYGEN)
#include "mk_raygen.cl"
#elif defined(REF_TU_MK_NEXT_VERTEX)
#include "mk_next_vertex.cl"
#elif defined(REF_TU_MK_SAMPLE_BSDF)
#include "mk_sample_bsdf.cl"
#elif defined(REF_TU_MK_SPLAT)
#include "mk_splat.cl"
#elif defined(REF_TU_MK_SPLAT_PREVIEW)
#include "mk_splat_preview.cl"
#else
#error "no REF_TU_* selected"
#endif
} // namespace

// REF_VARIANT (build_ref.py: -DREF_VARIANT=dn_ / bs_ / aos_ / lm_) goes into the exported names, so that the same kernel built
// with other reference build options (USE_OPTIX_DENOISER, USE_BITSTACK, without USE_SOA) or with the C library's math lives
// beside the base build in one library: ref_<variant><kernel>, ref_par_<variant><kernel>.
#ifndef REF_VARIANT
#define REF_VARIANT
#endif
#define REF_CAT3_(a, b, c) a##b##c
#define REF_CAT3(a, b, c) REF_CAT3_(a, b, c)
#ifdef SHIM_PARALLEL
#define REF_LOOP _Pragma("omp parallel for schedule(dynamic, 4096)") for (long long g = (long long)begin; g < (long long)end; ++g)
#define REF_NAME(n) REF_CAT3(ref_par_, REF_VARIANT, n)
#else
#define REF_LOOP for (long long g = (long long)begin; g < (long long)end; ++g)
#define REF_NAME(n) REF_CAT3(ref_, REF_VARIANT, n)
#endif

#define T(b) ((GPUTaskState *)(b)->tasks)
#define QL(b) ((QueueCounters *)(b)->queueLens)
#define P(b) ((RenderParams *)(b)->params)
#define MATARGS(b, q) T(b), QL(b), (b)->q, (b)->extensionQueue, (Material *)(b)->materials, (b)->texData, (TexDescriptor *)(b)->textures, P(b), (b)->numTasks

// microkernel integrator (src/mk_*.cl).  reset / splat / splatPreview are launched as a 2-D NDRange(width, height)
// (clcontext.cpp:712,742,748): work-item g of the flattened range is (g % width, g / width); the others are 1-D.
#define REF_2D(g) g_shim_gid = (size_t)(g) % P(b)->width; g_shim_gid1 = (size_t)(g) / P(b)->width
#define REF_1D(g) g_shim_gid = (size_t)(g); g_shim_gid1 = 0
#define ST(b) ((RenderStats *)(b)->stats)

extern "C"
{
#if defined(REF_TU_RESET)
    void REF_NAME(reset)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { g_shim_gid = (size_t)g; reset(T(b), b->pixels, b->denoiserAlbedo, b->denoiserNormal, QL(b), b->raygenQueue, P(b), b->numTasks); }
    }
#ifndef SHIM_PARALLEL
    // layout facts the rest of the repo relies on, checked by tests
    void REF_NAME(layout)(uint32_t out[8])
    {
        out[0] = sizeof(GPUTaskState); out[1] = sizeof(GPUNode); out[2] = sizeof(Triangle); out[3] = sizeof(Material);
        out[4] = sizeof(RenderParams); out[5] = sizeof(QueueCounters); out[6] = sizeof(TexDescriptor); out[7] = sizeof(Hit);
    }
#endif
#elif defined(REF_TU_RAYGEN)
    void REF_NAME(raygen)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { g_shim_gid = (size_t)g; genRays(T(b), P(b), QL(b), b->raygenQueue, b->extensionQueue, b->currPixelIdx, b->numTasks); }
    }
#elif defined(REF_TU_EXT)
    void REF_NAME(ext)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { g_shim_gid = (size_t)g; traceExtension(T(b), QL(b), b->extensionQueue, (Triangle *)b->tris, (GPUNode *)b->nodes, b->indices, P(b), b->numTasks); }
    }
#elif defined(REF_TU_SHADOW)
    void REF_NAME(shadow)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { g_shim_gid = (size_t)g; traceShadow(T(b), QL(b), b->shadowQueue, (Triangle *)b->tris, (GPUNode *)b->nodes, b->indices, P(b), b->numTasks); }
    }
#elif defined(REF_TU_LOGIC_SINGLE) || defined(REF_TU_LOGIC_SEPARATE)
#if defined(REF_TU_LOGIC_SINGLE)
    void REF_NAME(logic_single)(const RefBufs *b, size_t begin, size_t end)
#else
    void REF_NAME(logic_separate)(const RefBufs *b, size_t begin, size_t end)
#endif
    {
        const shim_image img = {b->envW, b->envH, b->envRGBA};
        REF_LOOP
        {
            g_shim_gid = (size_t)g;
            logic(T(b), b->pixels, b->denoiserNormal, b->denoiserAlbedo, QL(b), b->extensionQueue, b->shadowQueue, b->raygenQueue,
                  b->diffuseQueue, b->glossyQueue, b->ggxReflQueue, b->ggxRefrQueue, b->deltaQueue, (Triangle *)b->tris,
                  (GPUNode *)b->nodes, b->indices, &img, b->probTable, b->aliasTable, b->pdfTable, (Material *)b->materials,
                  b->texData, (TexDescriptor *)b->textures, P(b), b->numTasks, b->firstIteration);
        }
    }
#elif defined(REF_TU_MAT_ALL)
    void REF_NAME(mat_all)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontAllMaterials(MATARGS(b, diffuseQueue)); } }
#elif defined(REF_TU_MAT_DIFFUSE)
    void REF_NAME(mat_diffuse)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontDiffuse(MATARGS(b, diffuseQueue)); } }
#elif defined(REF_TU_MAT_GLOSSY)
    void REF_NAME(mat_glossy)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontGlossy(MATARGS(b, glossyQueue)); } }
#elif defined(REF_TU_MAT_GGX_REFL)
    void REF_NAME(mat_ggx_refl)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontGGXReflection(MATARGS(b, ggxReflQueue)); } }
#elif defined(REF_TU_MAT_GGX_REFR)
    void REF_NAME(mat_ggx_refr)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontGGXRefraction(MATARGS(b, ggxRefrQueue)); } }
#elif defined(REF_TU_POSTPROCESS)
    void REF_NAME(postprocess)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { g_shim_gid = (size_t)g; process(b->pixels, b->denoiserAlbedo, b->denoiserNormal, b->pixelsPreview, b->denoiserAlbedoGL, b->denoiserNormalGL, P(b), b->numTasks); }
    }
#elif defined(REF_TU_MK_RESET)
    void REF_NAME(mk_reset)(const RefBufs *b, size_t begin, size_t end)
    {
        REF_LOOP { REF_2D(g); reset(T(b), b->pixels, b->denoiserAlbedo, b->denoiserNormal, P(b), b->numTasks); }
    }
#elif defined(REF_TU_MK_RAYGEN)
    void REF_NAME(mk_raygen)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { REF_1D(g); genCameraRays(T(b), P(b), b->numTasks); } }
#elif defined(REF_TU_MK_NEXT_VERTEX)
    void REF_NAME(mk_next_vertex)(const RefBufs *b, size_t begin, size_t end)
    {
        const shim_image img = {b->envW, b->envH, b->envRGBA};
        REF_LOOP
        {
            REF_1D(g);
            nextVertex(T(b), (Material *)b->materials, b->texData, (TexDescriptor *)b->textures, b->denoiserNormal, (Triangle *)b->tris, (GPUNode *)b->nodes,
                       b->indices, P(b), ST(b), &img, b->pdfTable, b->numTasks);
        }
    }
#elif defined(REF_TU_MK_SAMPLE_BSDF)
    void REF_NAME(mk_sample_bsdf)(const RefBufs *b, size_t begin, size_t end)
    {
        const shim_image img = {b->envW, b->envH, b->envRGBA};
        REF_LOOP
        {
            REF_1D(g);
            sampleBsdf(T(b), b->denoiserAlbedo, (Material *)b->materials, b->texData, (TexDescriptor *)b->textures, &img, b->probTable, b->aliasTable,
                       b->pdfTable, (Triangle *)b->tris, (GPUNode *)b->nodes, b->indices, P(b), ST(b), b->numTasks);
        }
    }
#elif defined(REF_TU_MK_SPLAT)
    void REF_NAME(mk_splat)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { REF_2D(g); splat(T(b), b->pixels, P(b), ST(b), b->numTasks); } }
#elif defined(REF_TU_MK_SPLAT_PREVIEW)
    void REF_NAME(mk_splat_preview)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { REF_2D(g); splatPreview(T(b), b->pixels, P(b), b->numTasks); } }
#elif defined(REF_TU_MAT_DELTA)
    void REF_NAME(mat_delta)(const RefBufs *b, size_t begin, size_t end) { REF_LOOP { g_shim_gid = (size_t)g; wavefrontDelta(MATARGS(b, deltaQueue)); } }
#endif
}
