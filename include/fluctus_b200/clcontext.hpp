// clcontext.hpp -- C++ host side above the C ABI: a class with the method set of the reference's CLContext for the
// wavefront path (reference: src/clcontext.hpp:26-211), so the reference's Tracer::update()/runBenchmark() code
// (src/tracer.cpp:222-266, 431-470) compiles against it with the call sites unchanged.
//
//   * same method names and argument meaning; `const RenderParams&` arguments are accepted and ignored exactly like in the
//     reference, whose kernels read the device copy written by updateParams();
//   * same error behaviour: every failure throws std::runtime_error with the library's message (the reference throws from
//     clt::check, ext/CLT/src/utils.cpp:22-29); nothing calls exit();
//   * same asynchrony: enqueue* return immediately, finishQueue() is the only synchronisation, the QueueCounters handed to
//     enqueueGetCounters() is valid after finishQueue() (reference: CL_FALSE read, src/clcontext.cpp:668-671).
//
// Layout-compatible types: flx_RenderParams == RenderParams, flx_QueueCounters == QueueCounters, flx_Triangle ==
// RTTriangle/Triangle, flx_Node == Node/GPUNode, flx_Material == Material (src/geom.h, static_asserts in flx_api.cu).
// Header-only; link with -lfluctus_b200.
#pragma once

#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../fluctus_b200.h"

namespace fluctus_b200
{

typedef flx_RenderParams RenderParams;
typedef flx_QueueCounters QueueCounters;
typedef flx_RenderStats64 RenderStats64;
typedef flx_PerfNumbers PerfNumbers;

// What the reference passes as (BVH*, Scene*): the already-built arrays, borrowed for the duration of the call.
struct SceneArrays
{
    const flx_Triangle *tris = nullptr;
    uint32_t numTris = 0;
    const uint32_t *indices = nullptr;
    uint32_t numIndices = 0;
    const flx_Node *nodes = nullptr;
    uint32_t numNodes = 0;
    const flx_Material *materials = nullptr;
    uint32_t numMaterials = 0;
    const flx_TexDescriptor *textures = nullptr;
    uint32_t numTextures = 0;
    const uint8_t *texData = nullptr;
    size_t texBytes = 0;
};

struct EnvMapArrays // EnvironmentMap getters (reference: src/envmap.hpp:33-40)
{
    const float *rgb = nullptr;
    int width = 0, height = 0;
    const float *probTable = nullptr;
    const int32_t *aliasTable = nullptr;
    const float *pdfTable = nullptr;
};

class CLContext
{
  public:
    explicit CLContext(uint32_t numTasks = 1u << 20 /* wfBufferSize default, src/settings.cpp:20 */, int device = 0)
    {
        const int rc = flx_create(device, numTasks, &ctx);
        if (rc != 0)
            throw std::runtime_error(std::string("CLContext: ") + flx_last_error(nullptr));
    }
    ~CLContext() { flx_destroy(ctx); }
    CLContext(const CLContext &) = delete;
    CLContext &operator=(const CLContext &) = delete;

    // ---- setup (clcontext.hpp:62-79)
    void uploadSceneData(const SceneArrays &s)
    {
        verify(flx_upload_scene(ctx, s.tris, s.numTris, s.indices, s.numIndices, s.nodes, s.numNodes, s.materials, s.numMaterials, s.textures, s.numTextures,
                                s.texData, s.texBytes),
               "uploadSceneData");
    }
    void createEnvMap(const EnvMapArrays &m) { verify(flx_upload_envmap(ctx, m.rgb, m.width, m.height, m.probTable, m.aliasTable, m.pdfTable), "createEnvMap"); }
    void setupPixelStorage(uint32_t width, uint32_t height) { verify(flx_resize(ctx, width, height), "setupPixelStorage"); }
    void updateParams(const RenderParams &params) { verify(flx_update_params(ctx, &params), "updateParams"); }
    void recompileKernels(bool /*setArgs*/) {} // specialisations are selected from the params at launch (clcontext.cpp:852-874 has no analogue)

    // ---- wavefront stages (clcontext.hpp:43-48)
    void enqueueWfResetKernel(const RenderParams &) { verify(flx_enqueue_reset(ctx), "enqueueWfResetKernel"); }
    void enqueueWfRaygenKernel(const RenderParams &) { verify(flx_enqueue_raygen(ctx), "enqueueWfRaygenKernel"); }
    void enqueueWfExtRayKernel(const RenderParams &) { verify(flx_enqueue_extrays(ctx), "enqueueWfExtRayKernel"); }
    void enqueueWfShadowRayKernel(const RenderParams &) { verify(flx_enqueue_shadowrays(ctx), "enqueueWfShadowRayKernel"); }
    void enqueueWfLogicKernel(const RenderParams &, const bool firstIteration) { verify(flx_enqueue_logic(ctx, firstIteration ? 1 : 0), "enqueueWfLogicKernel"); }
    void enqueueWfMaterialKernels(const RenderParams &) { verify(flx_enqueue_materials(ctx), "enqueueWfMaterialKernels"); }

    void saveImage(const std::string &filename, const RenderParams &) { verify(flx_save_image(ctx, filename.c_str()), "saveImage"); } // clcontext.hpp:78

    void saveCheckpoint(const std::string &path) { verify(flx_checkpoint_save(ctx, path.c_str()), "saveCheckpoint"); }   // new: resumable renders
    void loadCheckpoint(const std::string &path) { verify(flx_checkpoint_load(ctx, path.c_str()), "loadCheckpoint"); }

    // ---- microkernel integrator (clcontext.hpp:35-40; the one Tracer::renderSingle uses, tracer.cpp:95-169)
    void enqueueResetKernel(const RenderParams &) { verify(flx_enqueue_mk_reset(ctx), "enqueueResetKernel"); }
    void enqueueRayGenKernel(const RenderParams &) { verify(flx_enqueue_mk_raygen(ctx), "enqueueRayGenKernel"); }
    void enqueueNextVertexKernel(const RenderParams &) { verify(flx_enqueue_mk_next_vertex(ctx), "enqueueNextVertexKernel"); }
    void enqueueBsdfSampleKernel(const RenderParams &) { verify(flx_enqueue_mk_sample_bsdf(ctx), "enqueueBsdfSampleKernel"); }
    void enqueueSplatKernel(const RenderParams &) { verify(flx_enqueue_mk_splat(ctx), "enqueueSplatKernel"); }
    void enqueueSplatPreviewKernel(const RenderParams &) { verify(flx_enqueue_mk_splat_preview(ctx), "enqueueSplatPreviewKernel"); }
    void renderSingleLoop(unsigned spp) { verify(flx_render_single(ctx, spp), "renderSingleLoop"); } // new: tracer.cpp:124-150 without host round trips

    void enqueuePostprocessKernel(const RenderParams &) { verify(flx_enqueue_postprocess(ctx), "enqueuePostprocessKernel"); } // clcontext.hpp:41
    std::vector<float> readPreview()
    {
        std::vector<float> rgba((size_t)flx_tile_pixels(ctx) * 4);
        verify(flx_read_preview(ctx, rgba.data(), rgba.size() / 4), "readPreview");
        return rgba;
    }

    // ---- queue bookkeeping (clcontext.hpp:53-57, 71)
    void enqueueClearWfQueues() { verify(flx_enqueue_clear_queues(ctx), "enqueueClearWfQueues"); }
    void enqueueGetCounters(QueueCounters *cnt) { verify(flx_enqueue_get_counters(ctx, cnt), "enqueueGetCounters"); }
    void finishQueue() { verify(flx_finish(ctx), "finishQueue"); }
    void updatePixelIndex(uint32_t numPixels, uint32_t numNewPaths) { verify(flx_update_pixel_index(ctx, numPixels, numNewPaths), "updatePixelIndex"); }
    void resetPixelIndex() { verify(flx_reset_pixel_index(ctx), "resetPixelIndex"); }
    uint32_t getNumTasks() const { return flx_num_tasks(ctx); }

    // ---- statistics (clcontext.hpp:66-70, 73)
    void resetStats() { verify(flx_reset_stats(ctx), "resetStats"); }
    RenderStats64 getStats()
    {
        RenderStats64 s;
        verify(flx_get_stats(ctx, &s), "getStats");
        return s;
    }
    void setProfiling(bool on) { verify(flx_set_profiling(ctx, on ? 1 : 0), "setProfiling"); }
    float kernelMilliseconds(int kernelId, uint32_t *launches = nullptr) // checkTracingPerf (clcontext.cpp:673-701) as a getter
    {
        float ms = 0.0f;
        verify(flx_get_kernel_ms(ctx, kernelId, &ms, launches), "checkTracingPerf");
        return ms;
    }

    // ---- new: fused loop, read-back, tiling, gather
    void render(uint32_t iterations) { verify(flx_render(ctx, iterations), "render"); }
    float renderTimed(uint32_t iterations)
    {
        float ms = 0.0f;
        verify(flx_render_timed(ctx, iterations, &ms), "renderTimed");
        return ms;
    }
    std::vector<float> readPixels()
    {
        std::vector<float> rgba((size_t)flx_tile_pixels(ctx) * 4);
        verify(flx_read_pixels(ctx, rgba.data(), rgba.size() / 4), "readPixels");
        return rgba;
    }
    void setTile(uint32_t part, uint32_t nParts, uint32_t stripeRows) { verify(flx_set_tile(ctx, part, nParts, stripeRows), "setTile"); }
    uint32_t tilePixels() const { return flx_tile_pixels(ctx); }
    void commInit(const void *uniqueId128, int rank, int nranks) { verify(flx_comm_init(ctx, uniqueId128, rank, nranks), "commInit"); }
    // asynchronous like the enqueue* calls (own stream, device-side snapshot of the frame as of this call); complete after
    // finishQueue(), or on return when a host destination is given
    void gatherPixels(int root, float *fullImageOrNull) { verify(flx_gather_pixels(ctx, root, fullImageOrNull), "gatherPixels"); }
    // root, after gatherPixels: the gathered frame -- accumulators, or their display pass (what saveImage writes on a tiled context's root)
    void readGathered(bool preview, float *rgba, size_t numPixels) { verify(flx_read_gathered(ctx, preview ? 1 : 0, rgba, numPixels), "readGathered"); }
    void readPixelsInto(float *rgba, size_t numPixels) { verify(flx_read_pixels(ctx, rgba, numPixels), "readPixels"); } // e.g. into hostAlloc()ed memory

    // ---- new: denoiser feature buffers (the reference's Tracer::useDenoiser / -DUSE_OPTIX_DENOISER build, src/kernel_impl.hpp:53)
    void setDenoiser(bool on) { verify(flx_set_denoiser(ctx, on ? 1 : 0), "setDenoiser"); }
    std::vector<float> readDenoiserAOV(bool albedo, bool processed)
    {
        std::vector<float> rgba((size_t)flx_tile_pixels(ctx) * 4);
        verify(flx_read_denoiser_aov(ctx, albedo ? 1 : 0, processed ? 1 : 0, rgba.data(), rgba.size() / 4), "readDenoiserAOV");
        return rgba;
    }

    // ---- new: hierarchy on the GPU instead of `new SBVH(&tris, ...)` (src/scene.cpp:574-590): the same Node[] / index arrays
    void buildBVH(const flx_Triangle *tris, uint32_t numTris, std::vector<flx_Node> &nodes, std::vector<uint32_t> &indices, int quality = FLX_BVH_PLOC_OPT, uint32_t maxLeaf = 8,
                  float *buildMs = nullptr)
    {
        nodes.resize(numTris ? 2 * (size_t)numTris - 1 : 0);
        indices.resize(numTris);
        uint32_t numNodes = 0;
        verify(flx_build_bvh(ctx, tris, numTris, maxLeaf, quality, nodes.data(), (uint32_t)nodes.size(), &numNodes, indices.data(), buildMs), "buildBVH");
        nodes.resize(numNodes);
    }

    // ---- new: page-locked host memory for the arrays handed to uploadSceneData / filled by readPixelsInto (DMA at link speed)
    static void *hostAlloc(size_t bytes)
    {
        void *p = nullptr;
        if (flx_host_alloc(&p, bytes) != 0)
            throw std::runtime_error(std::string("hostAlloc: ") + flx_last_error(nullptr));
        return p;
    }
    static void hostFree(void *p) { flx_host_free(p); }

    flx_ctx *handle() { return ctx; }

  private:
    void verify(int rc, const char *what) // CLContext::verify -> clt::check (clcontext.cpp:931-936)
    {
        if (rc != 0)
            throw std::runtime_error(std::string(what) + " failed (" + std::to_string(rc) + "): " + flx_last_error(ctx));
    }
    flx_ctx *ctx = nullptr;
};

} // namespace fluctus_b200
