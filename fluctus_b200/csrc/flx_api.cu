// flx_api.cu -- the C ABI of include/fluctus_b200.h: device memory, scene upload + BVH repack,
// stage launches on one in-order stream, counters/stats/timing, tiling and the NCCL gather.
// Mirrors the method set of the reference's CLContext (src/clcontext.hpp:26-211; implementation
// src/clcontext.cpp) -- each function below names the method it stands in for.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <queue>
#include <string>
#include <type_traits>
#include <vector>

#include "flx_bvh_build.cuh"
#include "flx_bvh_repack.cuh"
#include "flx_kernels.cuh"
#include "flx_mk.cuh"
#include "flx_trace_greedy.cuh"
#include "flx_trace_persistent.cuh"

static_assert(sizeof(flx_RenderParams) == 240, "RenderParams layout (geom.h:163-180)");
static_assert(sizeof(flx_AreaLight) == 96 && sizeof(flx_Camera) == 80, "AreaLight/Camera layout");
static_assert(sizeof(flx_Node) == 48 && sizeof(flx_Triangle) == 160 && sizeof(flx_Material) == 80, "scene layouts");
static_assert(sizeof(flx_TexDescriptor) == 12 && sizeof(flx_QueueCounters) == 32, "descriptor/counter layouts");
static_assert(offsetof(flx_RenderParams, camera) == 96 && offsetof(flx_RenderParams, width) == 184 && offsetof(flx_RenderParams, worldRadius) == 228, "RenderParams offsets");
static_assert(offsetof(flx_Node, nPrims) == 40 && offsetof(flx_Triangle, matId) == 144 && offsetof(flx_Material, type) == 68, "field offsets");

struct Id128 // ncclUniqueId is a 128-byte opaque struct passed by value
{
    char bytes[128];
};

namespace
{
thread_local std::string g_create_error;

struct EventPair
{
    cudaEvent_t a, b;
    int kernel;
};

// NCCL entry points resolved at run time (the process usually already holds torch's libnccl.so.2)
struct NcclApi
{
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, Id128, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
} // namespace

struct flx_ctx
{
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // flx_render runs the shadow-ray kernel here so its start overlaps the extension kernel's tail
    cudaStream_t cur = nullptr;       // stream the next traversal launch goes to (== stream except inside flx_render)
    cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaStream_t stream3 = nullptr;   // flx_render runs the display pass here, beside the traversal stages (which never touch the accumulator)
    cudaEvent_t evPostFork = nullptr, evPostJoin = nullptr;
    cudaEvent_t evPixels = nullptr;   // recorded after every launch that writes the accumulator (logic's splat, the resets, mk splat)
    bool pixelsEventValid = false;
    int overlapPostprocess = 1;
    int overlapTrace = 1;
    // Every ABI call that touches the stream bumps opSeq; flx_enqueue_extrays remembers its number.  A flx_enqueue_shadowrays
    // that comes DIRECTLY after it (the order of the reference's loop, tracer.cpp:253-254 / 437-438) may then run on the second
    // stream from the fork point recorded before the extension launch: the two stages have disjoint inputs and outputs, and the
    // main stream waits for the join before anything enqueued later, so the in-order semantics a caller sees are unchanged.
    unsigned long long opSeq = 0, extSeq = ~0ull;
    // Deferred stages of the per-stage ABI (fuseStages): flx_enqueue_logic only notes the request (1), a flx_enqueue_raygen that
    // follows directly is noted too (2), and a flx_enqueue_materials after that launches the fused kernel for all three --
    // the order of the reference's loop (tracer.cpp:247-251 / 433-435).  ANY other call first launches what is pending as the
    // separate kernels, so nothing a caller can observe differs from immediate execution.
    int pendingStages = 0, pendingFirstIteration = 0;
    int postprocessInLoop = 1;        // flx_render runs the display pass every iteration, like the reference's loop (tracer.cpp:447)
    uint32_t numTasks = 0;
    std::string error;

    // path state
    uint32_t *tasks = nullptr;
    uint32_t *queues[8] = {};
    flx_QueueCounters *counters = nullptr, *snapshot = nullptr;
    flx_RenderStats64 *stats = nullptr;
    uint32_t *currPixelIdx = nullptr;
    unsigned long long *scanTiles = nullptr;
    uint32_t *scanTicket = nullptr;
    uint32_t numScanTiles = 0;
    uint32_t hostPixelIdx = 0; // CLContext::pixelIdx (clcontext.hpp:164)
    bool pixelIdxAdvancedOnDevice = false; // flx_render advances the device copy (k_end_iteration); fetch before the host advances it again
    uint32_t *pinnedPixelIdx = nullptr;    // staging ring for the asynchronous 4-byte writes of flx_update_pixel_index
    int pixelIdxRingPos = 0;

    // pinned staging for asynchronous counter reads
    flx_QueueCounters *pinnedCounters = nullptr;
    std::vector<std::pair<int, flx_QueueCounters *>> pendingCounterReads; // (slot in pinned ring, host destination)
    static const int kCounterRing = 64;
    int counterRingPos = 0;

    // scene
    flx_Triangle *tris = nullptr;
    flx_Material *materials = nullptr;
    float4 *kdGamma = nullptr;
    flx_TexDescriptor *texDesc = nullptr;
    uint8_t *texData = nullptr;
    float4 *tnodes = nullptr, *ttris = nullptr, *tattr = nullptr; // traversal layout (flx_trace.cuh): inner nodes, leaf references, per-triangle attributes
    int rootRef = 0;
    uint32_t nTris = 0, nTNodes = 0, nTTris = 0, treeletNodes = 0;
    uint32_t materialTypes = 0;  // OR of the uploaded materials' type bits (Scene::getMaterialTypes, src/scene.cpp:25,299): kernels for
    bool otherTypes = false;     // BSDF types the scene does not contain are not launched; otherTypes: a type none of the lists knows
    bool sceneReady = false;
    size_t sceneBytes = 0;
    // Device allocations are kept and reused while they are large enough (capacity by owning pointer variable): re-uploading a scene
    // or re-creating the image at the same size costs no cudaMalloc / cudaFree -- which take tens of microseconds on a quiet box and
    // have been seen to take hundreds of MILLIseconds on a shared one (gpurun_out/call_r2_07.log: uploadSceneData 75 .. 686 ms).
    std::map<const void *, size_t> capacity;
    unsigned char *repackPool = nullptr; // scratch of the device repack (nodes, indices, scans), kept between uploads
    uint64_t sceneHash = 0;      // fingerprint of the uploaded scene (sizes, materials, a sample of nodes and triangles): checkpoints carry it

    // environment map
    float *envRGBA = nullptr, *probTable = nullptr, *pdfTable = nullptr;
    int32_t *aliasTable = nullptr;
    int envW = 1, envH = 1;

    // image
    float *pixels = nullptr, *denoiserAlbedo = nullptr, *denoiserNormal = nullptr, *preview = nullptr;
    uint8_t *dirtyPixels = nullptr;  // per pixel: accumulator changed since the display pass last saw it (k_postprocess)
    bool previewStale = true;        // the whole preview must be recomputed (new image, new exposure / tone-map operator, checkpoint)
    int dirtyPostprocess = 1;        // 0: recompute every pixel every pass
    uint32_t width = 0, height = 0, tilePixels = 0;
    uint32_t part = 0, nParts = 1, stripeRows = 1;
    float *gatherBuf = nullptr, *fullImage = nullptr; // rank-major gather target and de-interleaved full image (root)
    float *fullPreview = nullptr;                     // display pass of the full image (root; allocated on first use)
    size_t fullPreviewPixels = 0;
    int lastGatherRoot = -1;                          // root of the most recent flx_gather_pixels (-1: none yet)
    size_t gatherBufPixels = 0, fullImagePixels = 0;  // capacities, tracked separately (a resize can grow one and not the other)
    // The gather runs on its own stream from a SNAPSHOT of the accumulator (one device-to-device copy on the render stream, a
    // few microseconds), so the render stream goes on splatting into `pixels` while NCCL moves the snapshot: the collective is
    // off the critical path even when it runs every iteration.  evGatherDone guards the snapshot against the next gather.
    cudaStream_t gatherStream = nullptr;
    cudaEvent_t evSnapshot = nullptr, evGatherDone = nullptr; // evGatherDone: the most recent gather has finished
    cudaEvent_t evSnapshotFree[2] = {nullptr, nullptr};        // the gather that read snapshot buffer k has finished
    float *gatherSnapshot = nullptr;                           // TWO snapshot buffers, used alternately (double-buffered accumulator)
    size_t gatherSnapshotPixels = 0;
    int gatherParity = 0;
    bool gatherInFlight = false, snapshotBusy[2] = {false, false};
    unsigned char *imageBlock = nullptr; // pixels | denoiserAlbedo | denoiserNormal | preview | dirtyPixels in one allocation
    int denoiser = 0;                    // flx_set_denoiser: accumulate the feature buffers (reference: Tracer::useDenoiser -> -DUSE_OPTIX_DENOISER)
    float *aovOut = nullptr;             // display-pass outputs of the two feature buffers (normal | albedo), allocated on first use

    flx_RenderParams params;
    bool paramsSet = false;
    float tanHalfFov = 0.0f;
    int pinholeCamera = 0;    // the thin-lens offset is exactly zero for every sample (lens_offset, flx_kernels.cuh)

    // traversal work counters (flx_set_counting): [0..4] extension V,B,T,U,rays  [5..9] shadow V,B,T,U,rays
    unsigned long long *traceCounts = nullptr;
    bool counting = false;

    // tuning knobs (flx_set_tuning)
    int traceVariant = 1;     // 0: one ray per thread, 1: persistent threads + dynamic fetch, while-while phases (production), 2: 1 + top-of-tree
                              // treelet in shared memory, 3: persistent threads, one majority step per iteration (flx_trace_greedy.cuh; measured equal)
    int useMaterialMask = 1;     // fused logic kernel: compile-time lobe set chosen from materialTypes
    int bvhDepthLimit = 62;         // deepest PLOC tree flx_build_bvh hands out (tests lower it to reach the fallback)
    int bvhReinsertIterations = 16; // flx_build_bvh, FLX_BVH_PLOC_OPT: iterations of the parallel-reinsertion post-pass
    int bvhTriCostPercent = 100; // flx_build_bvh: SAH cost of a triangle test relative to a box test, in percent (reference constants: 100)
    int logicTile = 256;      // paths per tile (= threads per CTA) of the logic kernel: 256 or 128
    int gatherDirect = 0;     // flx_gather_pixels: 1 = receive every stripe straight into its rows of the full image; 0 (default) = rank-major
                              // buffer + de-interleave.  Measured (C5 on 2 GPUs, profiles/r2_gather_direct_2gpu.txt): NCCL's cost per point-to-point
                              // operation makes 135 stripe-sized receives take 1.8-2.0 ms against 0.44-0.55 ms for one receive + the pass
    int gatherPriority = 0;   // 1: the gather stream gets the render stream's (high) priority instead of the lowest
    int innerBias = 0;        // variant 3: an inner-node step runs when lanes-at-inner + bias >= lanes-at-triangle
    int topNodes = 2047;      // variant 2: treelet nodes staged per CTA (64 B each)
    int fetchThreshold = 16;  // refill when fewer lanes than this still hold a ray
    int extMinBlocks = 9, shadowMinBlocks = 10; // variant 1: resident 128-thread CTAs per SM the kernels are compiled for
    int maxL1 = 0;
    int smemStack = 0;        // variant 1: first 4 / 8 / 24 stack levels in shared memory ([level][thread], conflict-free); 0 = all in local memory
    int fetchChunk = 32;     // queue entries a warp reserves per atomic
    int repackOnHost = 0;     // flx_upload_scene: build the traversal layout with the host code instead of the device kernels (checker)
    int l2Persist = 0;        // 0 off; 1 / 2: persisting-L2 access window over the TTri / TNode array on the two traversal streams
    int logicMinBlocks = 3;   // resident 256-thread CTAs per SM the logic kernel is compiled for (register budget)
    int fuseStages = 1;       // flx_render: logic + raygen + materials as one kernel
    int fusedMinBlocks = 3;   // register budget of that kernel (1..4 resident CTAs of 256 per SM; 3 measured best)
    int innerMin = 8;         // leave the inner-node phase when fewer lanes than this are still at inner nodes
    int traceBlocksPerSM = 0; // 0: occupancy calculator
    int numSMs = 148;
    int maxDynSmem = 48 * 1024;
    int occExt[3] = {0, 0, 0}, occShadow[3] = {0, 0, 0}; // resident CTAs per SM of the persistent kernels (min-blocks 8, 9, 10), asked once
    uint32_t *fetchCounters = nullptr; // [0] extension, [1] shadow, [2] microkernel nextVertex, [3] microkernel light samples
    bool scanClean = false;        // the logic kernel's scan state was zeroed on the device by k_end_iteration (flx_render)
    uint32_t lastLogicTiles = 0;   // tiles the most recent logic launch used (= status words it left non-zero)
    bool fetchClean[2] = {false, false}; // [0] / [1] were zeroed on the device by the previous iteration's k_end_iteration (flx_render)

    // microkernel integrator (flx_mk.cuh), allocated on first use
    uint32_t *mkScratch = nullptr;  // MK_X_SLOTS x numTasks
    uint32_t *mkRayQueue = nullptr; // 2 x numTasks candidate shadow rays
    uint32_t *mkRayCount = nullptr;
    uint32_t *mkTypeQueues = nullptr; // MK_NUM_LISTS x numTasks: vertices to shade, by BSDF type
    uint32_t *mkTypeCounts = nullptr;

    // timing
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    bool profiling = false;
    std::vector<EventPair> pendingEvents, freeEvents;
    double kernelMs[FLX_K_COUNT] = {};
    uint32_t kernelLaunches[FLX_K_COUNT] = {};

    // NCCL
    NcclApi nccl;
    void *comm = nullptr;
    int rank = 0, nranks = 1;
};

namespace
{
int fail(flx_ctx *ctx, int code, const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx)
        ctx->error = buf;
    else
        g_create_error = buf;
    return code;
}

#define CU(call)                                                                                                       \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(ctx, (int)e_, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__);     \
    } while (0)

#define REQUIRE(cond, msg)                                                                                             \
    do                                                                                                                 \
    {                                                                                                                  \
        if (!(cond))                                                                                                   \
            return fail(ctx, FLX_E_INVALID, "%s", msg);                                                                \
    } while (0)

template <class T> void freeDev(T *&p)
{
    if (p)
        cudaFree(p);
    p = nullptr;
}

Frame makeFrame(const flx_ctx *c)
{
    Frame f;
    f.tasks.base = c->tasks;
    f.tasks.n = c->numTasks;
    f.counters = c->counters;
    for (int i = 0; i < 8; i++)
        f.queues[i] = c->queues[i];
    f.pixels = c->pixels;
    f.dirty = c->dirtyPixels;
    f.denoiserAlbedo = c->denoiserAlbedo;
    f.denoiserNormal = c->denoiserNormal;
    f.denoiser = c->denoiser;
    f.currPixelIdx = c->currPixelIdx;
    f.numTasks = c->numTasks;
    f.tilePixels = c->tilePixels;
    f.part = c->part;
    f.nParts = c->nParts;
    f.stripeRows = c->stripeRows;
    f.tanHalfFov = c->tanHalfFov;
    f.pinholeCamera = c->pinholeCamera;
    return f;
}

SceneView makeScene(const flx_ctx *c)
{
    SceneView s;
    s.tris = c->tris;
    s.materials = c->materials;
    s.textures = c->texDesc;
    s.texData = c->texData;
    s.kdGamma = c->kdGamma;
    s.envRGBA = c->envRGBA;
    s.envW = c->envW;
    s.envH = c->envH;
    s.probTable = c->probTable;
    s.aliasTable = c->aliasTable;
    s.pdfTable = c->pdfTable;
    return s;
}

BvhView makeBvh(const flx_ctx *c)
{
    BvhView b;
    b.nodes = c->tnodes;
    b.tris = c->ttris;
    b.attr = c->tattr;
    b.rootRef = c->rootRef;
    return b;
}

IterationState makeIter(const flx_ctx *c)
{
    IterationState it;
    it.counters = c->counters;
    it.snapshot = c->snapshot;
    it.fetch = nullptr;
    it.scanTiles = nullptr;
    it.nScanTiles = 0;
    it.scanTicket = nullptr;
    it.stats = c->stats;
    it.currPixelIdx = c->currPixelIdx;
    it.tilePixels = c->tilePixels;
    return it;
}

uint32_t localRows(uint32_t height, uint32_t part, uint32_t nParts, uint32_t stripeRows)
{
    uint32_t rows = 0;
    for (uint32_t s = part; s * stripeRows < height; s += nParts)
        rows += std::min(stripeRows, height - s * stripeRows);
    return rows;
}

// ---- timing helpers
struct Timed
{
    flx_ctx *c;
    int k;
    EventPair ev;
    bool on;
    Timed(flx_ctx *ctx, int kernel) : c(ctx), k(kernel), on(ctx->profiling)
    {
        if (!on)
            return;
        if (!c->freeEvents.empty())
        {
            ev = c->freeEvents.back();
            c->freeEvents.pop_back();
        }
        else
        {
            cudaEventCreate(&ev.a);
            cudaEventCreate(&ev.b);
        }
        ev.kernel = k;
        cudaEventRecord(ev.a, c->cur);
    }
    ~Timed()
    {
        c->kernelLaunches[k]++;
        if (!on)
            return;
        cudaEventRecord(ev.b, c->cur);
        c->pendingEvents.push_back(ev);
    }
};

void drainEvents(flx_ctx *c)
{
    std::vector<EventPair> notReady; // e.g. a gather still running on its own stream: asked again at the next drain
    for (auto &e : c->pendingEvents)
    {
        float ms = 0.0f;
        const cudaError_t r = cudaEventElapsedTime(&ms, e.a, e.b);
        if (r == cudaErrorNotReady)
        {
            notReady.push_back(e);
            continue;
        }
        if (r == cudaSuccess)
            c->kernelMs[e.kernel] += ms;
        c->freeEvents.push_back(e);
    }
    c->pendingEvents.swap(notReady);
}

int flushPending(flx_ctx *ctx);
// every ABI call that observes or changes device state passes through here (see flx_ctx::opSeq, flx_ctx::pendingStages)
int touch(flx_ctx *ctx, bool keepPending = false)
{
    ctx->opSeq++;
    return keepPending ? 0 : flushPending(ctx);
}
#define TOUCH(ctx)                                                                                                     \
    do                                                                                                                 \
    {                                                                                                                  \
        int rc_ = touch(ctx);                                                                                          \
        if (rc_)                                                                                                       \
            return rc_;                                                                                                \
    } while (0)

int checkReady(flx_ctx *ctx, bool needScene, bool needImage, bool keepPending = false)
{
    if (!ctx)
        return FLX_E_INVALID;
    if (int rcTouch = touch(ctx, keepPending))
        return rcTouch;
    if (!ctx->paramsSet)
        return fail(ctx, FLX_E_NOT_READY, "flx_update_params has not been called");
    if (needImage && !ctx->pixels)
        return fail(ctx, FLX_E_NOT_READY, "flx_resize has not been called");
    if (needScene && !ctx->sceneReady)
        return fail(ctx, FLX_E_NOT_READY, "flx_upload_scene has not been called");
    return 0;
}

int launchCheck(flx_ctx *ctx, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess)
        return fail(ctx, (int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
    return 0;
}

// after a launch that writes the accumulator: the display pass may start from here (flx_enqueue_postprocess)
void markPixelsWritten(flx_ctx *ctx)
{
    ctx->pixelsEventValid = cudaEventRecord(ctx->evPixels, ctx->stream) == cudaSuccess;
}

unsigned streamingGrid(uint32_t n) { return std::max(1u, std::min((n + FLX_BLOCK - 1) / FLX_BLOCK, 148u * 16u)); }

// Depth of the deepest leaf (root = 0).  Links only go forward in the reference's depth-first layout, so one pass from the front
// sees every parent before its children.  The traversal stack holds FLX_STACK_DEPTH entries (reference: uint stack[64],
// src/bvh.cl:240) and a ray at a leaf of depth d can have d entries pending, so a deeper tree would overrun it: the in-repo
// builders stop at 62, a caller's array or an imported cache file (flx_hierarchy_import) is checked here.  Assumes the links
// were range-checked (child > parent) by the caller.
uint32_t hierarchyDepth(const flx_Node *nodes, uint32_t nNodes)
{
    std::vector<uint16_t> depth(nNodes, 0);
    uint32_t deepest = 0;
    for (uint32_t i = 0; i < nNodes; i++)
    {
        if (nodes[i].nPrims != 0)
            continue;
        const uint32_t d = std::min<uint32_t>(depth[i] + 1u, 0xffffu);
        depth[i + 1] = (uint16_t)d;
        depth[nodes[i].iStartOrRightChild] = (uint16_t)d;
        deepest = std::max(deepest, d);
    }
    return deepest;
}

// every triangle's material index against the material count, on the device copy (reading 160-byte records on the host just for
// one int each costs more than the whole upload); error = first offending triangle + 1
__global__ void __launch_bounds__(256) k_validate_triangles(const flx_Triangle *tris, uint32_t nTris, uint32_t nMaterials, uint32_t *error)
{
    const uint32_t i = blockIdx.x * 256 + threadIdx.x;
    if (i < nTris && (tris[i].matId < 0 || (uint32_t)tris[i].matId >= nMaterials))
        atomicMin(error, i + 1u);
}

// ---- BVH repack: reference Node[] (48 B, DFS, left = self + 1) + indices + Triangle[] (160 B)
//      -> TNode[] (64 B, inner nodes only) + TTri[] (48 B per leaf reference). See flx_trace.cuh.
struct Repacked
{
    std::vector<float4> nodes, tris;
    int rootRef = 0;
    uint32_t treeletNodes = 0;
};

int repackBvh(flx_ctx *ctx, const flx_Triangle *tris, uint32_t nTris, const uint32_t *indices, uint32_t nIndices, const flx_Node *nodes, uint32_t nNodes,
              Repacked &out)
{
    std::vector<int> ref(nNodes, 0); // child reference of every reference node
    // pass 1: validate, lay out the leaves' triangles in DFS order
    uint32_t nInner = 0;
    size_t nLeafTris = 0;
    for (uint32_t i = 0; i < nNodes; i++)
    {
        if (nodes[i].nPrims == 0)
        {
            const uint32_t r = nodes[i].iStartOrRightChild;
            if (i + 1 >= nNodes || r >= nNodes || r <= i + 1)
                return fail(ctx, FLX_E_INVALID, "node %u: child links out of range (left %u, right %u, %u nodes)", i, i + 1, r, nNodes);
            nInner++;
        }
        else
        {
            const uint32_t s = nodes[i].iStartOrRightChild, n = nodes[i].nPrims;
            if ((size_t)s + n > nIndices)
                return fail(ctx, FLX_E_INVALID, "leaf %u: index range [%u,%u) exceeds %u indices", i, s, s + n, nIndices);
            if (nLeafTris > 0x7ffffff0u)
                return fail(ctx, FLX_E_INVALID, "too many leaf references");
            ref[i] = ~(int)nLeafTris;
            nLeafTris += n;
        }
    }
    if (const uint32_t deep = hierarchyDepth(nodes, nNodes); deep > FLX_STACK_DEPTH)
        return fail(ctx, FLX_E_INVALID, "hierarchy is %u levels deep; the traversal stack holds %d entries (reference: uint stack[64], bvh.cl:240)", deep, FLX_STACK_DEPTH);
    // pass 2: number the inner nodes -- first a treelet grown from the root by always taking the pending node with the
    // largest box area (the nodes a random ray is most likely to visit; any prefix of this order is a connected
    // top-of-tree, which is what the TOP traversal variant stages in shared memory), then everything else in DFS order.
    const uint32_t kTreeletMax = 4096;
    std::vector<char> placed(nNodes, 0);
    uint32_t next = 0;
    if (nodes[0].nPrims == 0)
    {
        auto area = [&](uint32_t i) {
            const float dx = nodes[i].bmax.x - nodes[i].bmin.x, dy = nodes[i].bmax.y - nodes[i].bmin.y, dz = nodes[i].bmax.z - nodes[i].bmin.z;
            return dx * dy + dx * dz + dy * dz;
        };
        typedef std::pair<float, uint32_t> Item; // (area, ~index): larger area first, lower index first on ties
        std::priority_queue<Item> pq;
        pq.push(Item(area(0), ~0u));
        while (!pq.empty() && next < kTreeletMax)
        {
            const uint32_t i = ~pq.top().second;
            pq.pop();
            ref[i] = (int)next++;
            placed[i] = 1;
            const uint32_t kids[2] = {i + 1, nodes[i].iStartOrRightChild};
            for (uint32_t c : kids)
                if (nodes[c].nPrims == 0)
                    pq.push(Item(area(c), ~c));
        }
    }
    out.treeletNodes = next;
    for (uint32_t i = 0; i < nNodes; i++)
        if (nodes[i].nPrims == 0 && !placed[i])
            ref[i] = (int)next++;
    out.nodes.assign((size_t)nInner * 4, make_float4(0, 0, 0, 0));
    out.tris.assign(nLeafTris * 4, make_float4(0, 0, 0, 0));
    out.rootRef = ref[0];
    for (uint32_t i = 0; i < nNodes; i++)
    {
        if (nodes[i].nPrims == 0)
        {
            const flx_Node &L = nodes[i + 1], &R = nodes[nodes[i].iStartOrRightChild];
            float4 *q = &out.nodes[(size_t)ref[i] * 4];
            q[0] = make_float4(L.bmin.x, L.bmin.y, L.bmin.z, L.bmax.x);
            q[1] = make_float4(L.bmax.y, L.bmax.z, R.bmin.x, R.bmin.y);
            q[2] = make_float4(R.bmin.z, R.bmax.x, R.bmax.y, R.bmax.z);
            int4 links = make_int4(ref[i + 1], ref[nodes[i].iStartOrRightChild], 0, 0);
            memcpy(&q[3], &links, sizeof links);
        }
        else
        {
            const uint32_t s = nodes[i].iStartOrRightChild, n = nodes[i].nPrims;
            float4 *q = &out.tris[(size_t)(~ref[i]) * 4];
            for (uint32_t k = 0; k < n; k++, q += 4)
            {
                const uint32_t ti = indices[s + k];
                if (ti >= nTris || ti > 0x7fffffffu)
                    return fail(ctx, FLX_E_INVALID, "leaf %u references triangle %u of %u", i, ti, nTris);
                const flx_Triangle &T = tris[ti];
                uint32_t tag = ti | (k + 1 == n ? 0x80000000u : 0u);
                float tagf;
                memcpy(&tagf, &tag, 4);
                // the float differences below are the ones the reference forms per test (intersect.cl:66-67)
                q[0] = make_float4(T.v0.p.x, T.v0.p.y, T.v0.p.z, tagf);
                q[1] = make_float4(T.v1.p.x - T.v0.p.x, T.v1.p.y - T.v0.p.y, T.v1.p.z - T.v0.p.z, 0.0f);
                q[2] = make_float4(T.v2.p.x - T.v0.p.x, T.v2.p.y - T.v0.p.y, T.v2.p.z - T.v0.p.z, 0.0f);
            }
        }
    }
    return 0;
}

// The treelet of repackBvh (pass 2) on its own: the inner nodes a random ray is most likely to visit, grown from the root by
// always taking the pending node with the largest box area; returned in placement order.
std::vector<uint32_t> pickTreelet(const flx_Node *nodes, uint32_t nNodes)
{
    std::vector<uint32_t> order;
    const uint32_t kTreeletMax = 4096;
    if (nodes[0].nPrims != 0)
        return order;
    auto area = [&](uint32_t i) {
        const float dx = nodes[i].bmax.x - nodes[i].bmin.x, dy = nodes[i].bmax.y - nodes[i].bmin.y, dz = nodes[i].bmax.z - nodes[i].bmin.z;
        return dx * dy + dx * dz + dy * dz;
    };
    typedef std::pair<float, uint32_t> Item; // (area, ~index): larger area first, lower index first on ties
    std::priority_queue<Item> pq;
    pq.push(Item(area(0), ~0u));
    while (!pq.empty() && order.size() < kTreeletMax)
    {
        const uint32_t i = ~pq.top().second;
        pq.pop();
        order.push_back(i);
        const uint32_t kids[2] = {i + 1, nodes[i].iStartOrRightChild};
        for (uint32_t c : kids)
            if (c < nNodes && nodes[c].nPrims == 0)
                pq.push(Item(area(c), ~c));
    }
    return order;
}

// dst = a device block of at least `bytes`: the one it already holds when that is large enough, else a fresh one
template <class T> int reserveDev(flx_ctx *ctx, T *&dst, size_t bytes)
{
    size_t &cap = ctx->capacity[static_cast<const void *>(&dst)];
    if (dst && cap >= bytes)
        return 0;
    freeDev(dst);
    cap = 0;
    CU(cudaMalloc(&dst, bytes));
    cap = bytes;
    return 0;
}

template <class T> int uploadArray(flx_ctx *ctx, T *&dst, const T *src, size_t count, size_t minCount = 1)
{
    const size_t bytes = std::max(count, minCount) * sizeof(T);
    if (int rcReserve = reserveDev(ctx, dst, bytes))
        return rcReserve;
    // On the context's stream (the work streams are cudaStreamNonBlocking: the legacy default stream does not order against
    // them).  A pinned source (flx_host_alloc) goes by DMA at link speed; a pageable one is staged by the driver before the call
    // returns.  Callers synchronise the stream before they return, so the source is borrowed for the call only.
    if (bytes > count * sizeof(T))
        CU(cudaMemsetAsync(reinterpret_cast<unsigned char *>(dst) + count * sizeof(T), 0, bytes - count * sizeof(T), ctx->stream));
    if (count)
        CU(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    ctx->sceneBytes += bytes;
    return 0;
}

// flx_bvh_repack.cuh driven from the host: nodes / indices go up as they are, the traversal layout is made on the device
int repackOnDevice(flx_ctx *ctx, const uint32_t *indices, uint32_t nIndices, const flx_Node *nodes, uint32_t nNodes, uint32_t nTris, uint32_t nMaterials)
{
    const bool timing = std::getenv("FLX_DEBUG_TIMING") != nullptr;
    const auto tHost0 = std::chrono::steady_clock::now();
    // sizes and the treelet come from one cheap pass over the host nodes
    // ONE pass over the host nodes: link ranges, counts, and the depth of the deepest leaf (hierarchyDepth's recurrence inlined: links
    // only go forward, so a node's depth is known before its children are reached)
    size_t nInner = 0, nLeafTris = 0;
    std::vector<uint8_t> depth(nNodes, 0);
    uint32_t deep = 0;
    for (uint32_t i = 0; i < nNodes; i++)
    {
        if (nodes[i].nPrims == 0)
        {
            const uint32_t r = nodes[i].iStartOrRightChild;
            if (i + 1 >= nNodes || r >= nNodes || r <= i + 1)
                return fail(ctx, FLX_E_INVALID, "node %u: child links out of range (left %u, right %u, %u nodes)", i, i + 1, r, nNodes);
            nInner++;
            const uint32_t d = std::min<uint32_t>(depth[i] + 1u, 255u);
            depth[i + 1] = depth[r] = (uint8_t)d;
            deep = std::max(deep, d);
        }
        else
            nLeafTris += nodes[i].nPrims;
    }
    if (nLeafTris > 0x7ffffff0u)
        return fail(ctx, FLX_E_INVALID, "too many leaf references");
    if (deep > FLX_STACK_DEPTH)
        return fail(ctx, FLX_E_INVALID, "hierarchy is %u levels deep; the traversal stack holds %d entries (reference: uint stack[64], bvh.cl:240)", deep, FLX_STACK_DEPTH);
    const std::vector<uint32_t> treelet = pickTreelet(nodes, nNodes);
    if (timing)
        std::fprintf(stderr, "  flx_upload_scene: %-28s %8.3f ms\n", "host pass over the nodes",
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tHost0).count());

    // one allocation for all temporaries (cudaMalloc / cudaFree are the expensive part of a small job like this)
    auto align = [](size_t b) { return (b + 255) & ~(size_t)255; };
    const size_t szNodes = align((size_t)nNodes * sizeof(flx_Node)), szIdx = align((size_t)nIndices * 4), szTreelet = align(std::max<size_t>(treelet.size(), 1) * 4),
                 szPerNode = align((size_t)nNodes * 4);
    size_t tempBytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tempBytes, (uint32_t *)nullptr, (uint32_t *)nullptr, (int)nNodes, ctx->stream);
    const size_t szScan = align(std::max<size_t>(tempBytes, 16));
    int rc = 0;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == 0)
            rc = fail(ctx, (int)e, "flx_upload_scene: %s failed: %s", what, cudaGetErrorString(e));
    };
    cudaStream_t st = ctx->stream;
    if (int rcPool = reserveDev(ctx, ctx->repackPool, szNodes + szIdx + szTreelet + 5 * szPerNode + szScan + 256))
        return rcPool;
    unsigned char *pool = ctx->repackPool;
    auto cleanup = [&]() {};
    unsigned char *cursor = pool;
    auto take = [&](size_t bytes) { unsigned char *p = cursor; cursor += bytes; return p; };
    flx_Node *dNodes = reinterpret_cast<flx_Node *>(take(szNodes));
    uint32_t *dIndices = reinterpret_cast<uint32_t *>(take(szIdx)), *dTreelet = reinterpret_cast<uint32_t *>(take(szTreelet));
    int *dPos = reinterpret_cast<int *>(take(szPerNode));
    uint32_t *dFlag = reinterpret_cast<uint32_t *>(take(szPerNode)), *dPrims = reinterpret_cast<uint32_t *>(take(szPerNode));
    uint32_t *dScanI = reinterpret_cast<uint32_t *>(take(szPerNode)), *dScanL = reinterpret_cast<uint32_t *>(take(szPerNode));
    void *dTemp = take(szScan);
    uint32_t *dError = reinterpret_cast<uint32_t *>(take(256)); // [0] hierarchy error, [1] first triangle with a bad material + 1
    const size_t nodeBytes = std::max<size_t>(nInner, 1) * 64, triBytes = std::max<size_t>(nLeafTris, 1) * 64;
    if (int rcT = reserveDev(ctx, ctx->tnodes, nodeBytes))
        return rcT;
    if (int rcT = reserveDev(ctx, ctx->ttris, triBytes))
        return rcT;
    if (rc == 0)
    {
        cu(cudaMemcpyAsync(dNodes, nodes, (size_t)nNodes * sizeof(flx_Node), cudaMemcpyHostToDevice, st), "node upload");
        cu(cudaMemcpyAsync(dIndices, indices, (size_t)nIndices * 4, cudaMemcpyHostToDevice, st), "index upload");
        if (!treelet.empty())
            cu(cudaMemcpyAsync(dTreelet, treelet.data(), treelet.size() * 4, cudaMemcpyHostToDevice, st), "treelet upload");
        cu(cudaMemsetAsync(dPos, 0xff, (size_t)nNodes * 4, st), "memset");
        cu(cudaMemsetAsync(dError, 0, 4, st), "memset");
        cu(cudaMemsetAsync(dError + 1, 0xff, 4, st), "memset");
        if (nInner == 0)
            cu(cudaMemsetAsync(ctx->tnodes, 0, nodeBytes, st), "memset"); // every record is written by k_repack_emit otherwise
    }
    if (rc == 0)
    {
        RepackView r;
        r.nodes = dNodes; r.indices = dIndices; r.tris = ctx->tris;
        r.nNodes = nNodes; r.nIndices = nIndices; r.nTris = nTris; r.treeletCount = (uint32_t)treelet.size();
        r.treeletPos = dPos; r.flagInner = dFlag; r.leafPrims = dPrims; r.scanInner = dScanI; r.scanLeaf = dScanL;
        r.tnodes = ctx->tnodes; r.ttris = ctx->ttris; r.error = dError;
        const unsigned grid = (nNodes + 255) / 256;
        if (!treelet.empty())
            k_repack_scatter_treelet<<<(unsigned)((treelet.size() + 255) / 256), 256, 0, st>>>(dTreelet, (uint32_t)treelet.size(), dPos);
        k_repack_flags<<<grid, 256, 0, st>>>(r);
        cu(cub::DeviceScan::ExclusiveSum(dTemp, tempBytes, dFlag, dScanI, (int)nNodes, st), "scan");
        cu(cub::DeviceScan::ExclusiveSum(dTemp, tempBytes, dPrims, dScanL, (int)nNodes, st), "scan");
        k_repack_emit<<<grid, 256, 0, st>>>(r);
        k_validate_triangles<<<(nTris + 255) / 256, 256, 0, st>>>(ctx->tris, nTris, nMaterials, dError + 1);
        cu(cudaGetLastError(), "repack launch");
        uint32_t errs[2] = {0, 0xffffffffu};
        cu(cudaMemcpyAsync(errs, dError, 8, cudaMemcpyDeviceToHost, st), "error read-back");
        cu(cudaStreamSynchronize(st), "repack");
        const uint32_t err = errs[0];
        if (rc == 0 && errs[1] != 0xffffffffu)
            rc = fail(ctx, FLX_E_INVALID, "triangle %u has a material index outside the %u uploaded", errs[1] - 1u, nMaterials);
        if (rc == 0 && err)
        {
            const uint32_t kind = err >> 28, node = (err & 0x0fffffffu) - 1u;
            rc = kind == 1 ? fail(ctx, FLX_E_INVALID, "node %u: child links out of range", node)
                 : kind == 2 ? fail(ctx, FLX_E_INVALID, "leaf %u: index range exceeds %u indices", node, nIndices)
                             : fail(ctx, FLX_E_INVALID, "leaf %u references triangle(s) outside the %u uploaded", node, nTris);
        }
    }
    cleanup();
    if (rc)
        return rc;
    ctx->sceneBytes += nodeBytes + triBytes;
    ctx->rootRef = nodes[0].nPrims == 0 ? 0 : ~0; // the root is the first node the treelet places; a single-leaf scene starts at TTri 0
    ctx->nTNodes = (uint32_t)nInner;
    ctx->nTTris = (uint32_t)nLeafTris;
    ctx->treeletNodes = (uint32_t)treelet.size();
    return 0;
}

int loadNccl(flx_ctx *ctx)
{
    NcclApi &n = ctx->nccl;
    if (n.lib)
        return 0;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names)
    {
        n.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (n.lib)
            break;
    }
    if (!n.lib)
        return fail(ctx, FLX_E_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define SYM(field, name)                                                                                              \
    *(void **)(&n.field) = dlsym(n.lib, name);                                                                         \
    if (!n.field)                                                                                                      \
        return fail(ctx, FLX_E_NCCL, "libnccl lacks %s", name);
    SYM(GetUniqueId, "ncclGetUniqueId")
    SYM(CommInitRank, "ncclCommInitRank")
    SYM(CommDestroy, "ncclCommDestroy")
    SYM(AllGather, "ncclAllGather")
    SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd")
    SYM(Send, "ncclSend")
    SYM(Recv, "ncclRecv")
    SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return 0;
}

// rank-major gathered tiles -> full image rows (inverse of local_pixel_to_xy)
__global__ void k_deinterleave(const float4 *gathered, float4 *full, uint32_t width, uint32_t height, uint32_t nParts, uint32_t stripeRows, uint32_t maxTilePixels)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= width * height)
        return;
    const uint32_t x = i % width, y = i / width;
    const uint32_t stripe = y / stripeRows, within = y % stripeRows;
    const uint32_t part = stripe % nParts, localStripe = stripe / nParts;
    const uint32_t ly = localStripe * stripeRows + within;
    full[i] = gathered[(size_t)part * maxTilePixels + (size_t)ly * width + x];
}
} // namespace

template <bool ANYHIT, class COUNT, int MINB, int SDEPTH> static int launchPersistentV1(flx_ctx *ctx, uint32_t *fetch, unsigned long long *counts)
{
    auto kern = k_trace_persistent<ANYHIT, COUNT, FLX_TRACE_BLOCK, false, MINB, SDEPTH>;
    const size_t smem = (size_t)(SDEPTH > 0 ? SDEPTH : 0) * FLX_TRACE_BLOCK * sizeof(int);
    if (smem > 48 * 1024)
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (SDEPTH <= 0 && ctx->maxL1) // the L1 is what this kernel lives on: ask for the largest L1 carve-out
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxL1));
    int perSM = ctx->traceBlocksPerSM;
    // the production instantiation (no counters, local-memory stack) asks the occupancy calculator once per context
    int *cached = (std::is_same<COUNT, NoCount>::value && SDEPTH == 0 && !ctx->maxL1) ? &(ANYHIT ? ctx->occShadow : ctx->occExt)[MINB - 8] : nullptr;
    if (perSM <= 0 && cached && *cached > 0)
        perSM = *cached;
    if (perSM <= 0)
    {
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, FLX_TRACE_BLOCK, smem));
        if (cached)
            *cached = perSM;
    }
    const unsigned grid = (unsigned)std::max(1, perSM) * (unsigned)ctx->numSMs;
    kern<<<grid, FLX_TRACE_BLOCK, smem, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, fetch, ctx->fetchThreshold, ctx->innerMin, ctx->fetchChunk, 0, counts, MkView{});
    return 0;
}

template <bool ANYHIT, class COUNT, int MINB> static int launchGreedy(flx_ctx *ctx, uint32_t *fetch, unsigned long long *counts)
{
    auto kern = k_trace_greedy<ANYHIT, COUNT, MINB>;
    int perSM = ctx->traceBlocksPerSM;
    if (perSM <= 0)
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, FLX_TRACE_BLOCK, 0));
    const unsigned grid = (unsigned)std::max(1, perSM) * (unsigned)ctx->numSMs;
    kern<<<grid, FLX_TRACE_BLOCK, 0, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, fetch, ctx->fetchThreshold, ctx->fetchChunk, ctx->innerBias, counts, MkView{});
    return 0;
}

template <bool ANYHIT, class COUNT> static int launchPersistentT(flx_ctx *ctx, uint32_t *fetch, unsigned long long *counts)
{
    if (ctx->traceVariant == 3)
    {
        switch (ANYHIT ? ctx->shadowMinBlocks : ctx->extMinBlocks)
        {
        case 8: return launchGreedy<ANYHIT, COUNT, 8>(ctx, fetch, counts);
        case 9: return launchGreedy<ANYHIT, COUNT, 9>(ctx, fetch, counts);
        default: return launchGreedy<ANYHIT, COUNT, 10>(ctx, fetch, counts);
        }
    }
    if (ctx->traceVariant == 2)
    {
        constexpr int BLOCK = 1024; // one persistent CTA per SM owns the staged treelet
        auto kern = k_trace_persistent<ANYHIT, COUNT, BLOCK, true, 1, 0>;
        const int top = (int)std::min<uint32_t>({(uint32_t)ctx->topNodes, ctx->treeletNodes, ctx->nTNodes, (uint32_t)((ctx->maxDynSmem - 1024) / 64)});
        const size_t smem = (size_t)top * 64;
        CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<ctx->numSMs, BLOCK, smem, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, fetch, ctx->fetchThreshold, ctx->innerMin, ctx->fetchChunk, top, counts, MkView{});
        return 0;
    }
    // register budget: the kernel is compiled for MINB resident 128-thread CTAs per SM; stack: local memory, or its first 4 / 8 / 24
    // levels in shared memory
    const int mb = ANYHIT ? ctx->shadowMinBlocks : ctx->extMinBlocks;
#define FLX_BY_MINB(SD)                                                                                                                                        \
    switch (mb)                                                                                                                                                \
    {                                                                                                                                                          \
    case 8: return launchPersistentV1<ANYHIT, COUNT, 8, SD>(ctx, fetch, counts);                                                                               \
    case 9: return launchPersistentV1<ANYHIT, COUNT, 9, SD>(ctx, fetch, counts);                                                                               \
    default: return launchPersistentV1<ANYHIT, COUNT, 10, SD>(ctx, fetch, counts);                                                                             \
    }
    switch (ctx->smemStack)
    {
    case -1: FLX_BY_MINB(-1)
    case 0: FLX_BY_MINB(0)
    case 4: FLX_BY_MINB(4)
    case 8: FLX_BY_MINB(8)
    default: FLX_BY_MINB(24)
    }
#undef FLX_BY_MINB
}

template <bool ANYHIT> static int launchPersistent(flx_ctx *ctx)
{
    uint32_t *fetch = ctx->fetchCounters + (ANYHIT ? 1 : 0);
    if (ctx->fetchClean[ANYHIT ? 1 : 0]) // flx_render: the previous iteration's last kernel left it at zero (one memset node less per launch)
        ctx->fetchClean[ANYHIT ? 1 : 0] = false;
    else
        CU(cudaMemsetAsync(fetch, 0, sizeof(uint32_t), ctx->cur));
    Timed tm(ctx, ANYHIT ? FLX_K_SHADOWRAYS : FLX_K_EXTRAYS);
    int rc = ctx->counting ? launchPersistentT<ANYHIT, RayCount>(ctx, fetch, ctx->traceCounts + (ANYHIT ? 5 : 0)) : launchPersistentT<ANYHIT, NoCount>(ctx, fetch, nullptr);
    if (rc)
        return rc;
    return launchCheck(ctx, ANYHIT ? "k_trace_persistent<shadow>" : "k_trace_persistent<extension>");
}

// ---- microkernel integrator plumbing
static int ensureMk(flx_ctx *ctx)
{
    if (ctx->mkScratch)
        return 0;
    CU(cudaMalloc(&ctx->mkScratch, (size_t)ctx->numTasks * MK_X_SLOTS * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->mkRayQueue, (size_t)ctx->numTasks * 2 * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->mkRayCount, sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->mkTypeQueues, (size_t)ctx->numTasks * MK_NUM_LISTS * sizeof(uint32_t)));
    CU(cudaMalloc(&ctx->mkTypeCounts, MK_NUM_LISTS * sizeof(uint32_t)));
    CU(cudaMemsetAsync(ctx->mkScratch, 0, (size_t)ctx->numTasks * MK_X_SLOTS * sizeof(uint32_t), ctx->stream));
    CU(cudaMemsetAsync(ctx->mkRayCount, 0, sizeof(uint32_t), ctx->stream));
    return 0;
}

// L2 residency hint for the traversal working set: the repacked hierarchy fits the 126 MB L2 but shares it with the streamed path
// state.  One access-policy window per stream, so one array is covered: 1 = TTri (the larger, more randomly read), 2 = TNode.
static int applyL2Persist(flx_ctx *ctx)
{
    if (!ctx->sceneReady)
        return 0;
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    if (ctx->l2Persist)
    {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, ctx->device));
        const void *base = ctx->l2Persist == 1 ? (const void *)ctx->ttris : (const void *)ctx->tnodes;
        size_t bytes = (ctx->l2Persist == 1 ? (size_t)ctx->nTTris : (size_t)ctx->nTNodes) * 64;
        bytes = std::min(bytes, (size_t)prop.accessPolicyMaxWindowSize);
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min(bytes, (size_t)prop.persistingL2CacheMaxSize)));
        attr.accessPolicyWindow.base_ptr = const_cast<void *>(base);
        attr.accessPolicyWindow.num_bytes = bytes;
        attr.accessPolicyWindow.hitRatio = 1.0f;
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    else
    {
        attr.accessPolicyWindow.num_bytes = 0; // disables the window
        attr.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        attr.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    CU(cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    CU(cudaStreamSetAttribute(ctx->stream2, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (!ctx->l2Persist)
        CU(cudaCtxResetPersistingL2Cache());
    return 0;
}

static MkView makeMk(const flx_ctx *c)
{
    MkView m;
    m.scratch.base = c->mkScratch;
    m.scratch.n = c->numTasks;
    m.rayQueue = c->mkRayQueue;
    m.rayCount = c->mkRayCount;
    m.typeQueues = c->mkTypeQueues;
    m.typeCounts = c->mkTypeCounts;
    m.limit = std::min(c->tilePixels, c->numTasks); // min(width * height, numTasks), e.g. mk_raygen.cl:9
    m.stats = c->stats;
    return m;
}

static unsigned mkGrid(uint32_t n) { return std::max(1u, (n + FLX_BLOCK - 1) / FLX_BLOCK); }

// one launch of the persistent traversal kernel in a microkernel mode (closest hit over the paths in phase
// MK_RT_NEXT_VERTEX, or any hit over the light-sample ray list)
template <bool ANYHIT> static int launchMkTrace(flx_ctx *ctx, const MkView &mk)
{
    constexpr int MODE = ANYHIT ? TRACE_MK_NEE : TRACE_MK_NEXT;
    constexpr int MINB = ANYHIT ? 10 : 9;
    uint32_t *fetch = ctx->fetchCounters + (ANYHIT ? 3 : 2);
    CU(cudaMemsetAsync(fetch, 0, sizeof(uint32_t), ctx->stream));
    int perSM = ctx->traceBlocksPerSM;
    if (ctx->traceVariant == 3)
    {
        auto kern = k_trace_greedy<ANYHIT, NoCount, MINB, MODE>;
        if (perSM <= 0)
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, FLX_TRACE_BLOCK, 0));
        const unsigned grid = (unsigned)std::max(1, perSM) * (unsigned)ctx->numSMs;
        kern<<<grid, FLX_TRACE_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, fetch, ctx->fetchThreshold, ctx->fetchChunk, ctx->innerBias, nullptr, mk);
        return launchCheck(ctx, ANYHIT ? "k_trace_greedy<mk light samples>" : "k_trace_greedy<mk nextVertex>");
    }
    auto kern = k_trace_persistent<ANYHIT, NoCount, FLX_TRACE_BLOCK, false, MINB, 0, MODE>;
    if (perSM <= 0)
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, FLX_TRACE_BLOCK, 0));
    const unsigned grid = (unsigned)std::max(1, perSM) * (unsigned)ctx->numSMs;
    kern<<<grid, FLX_TRACE_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, fetch, ctx->fetchThreshold, ctx->innerMin, ctx->fetchChunk, 0, nullptr, mk);
    return launchCheck(ctx, ANYHIT ? "k_trace_persistent<mk light samples>" : "k_trace_persistent<mk nextVertex>");
}

// With lazy module loading (the CUDA default) a kernel's code is loaded at its first launch, i.e. inside whatever the caller is
// timing.  The reference builds all its kernels in CLContext's constructor (clcontext.cpp:18-69); the analogue here is to touch
// every kernel of the default render loop once when the context is created.
static void preloadKernels()
{
    cudaFuncAttributes a;
#define PRELOAD(k) cudaFuncGetAttributes(&a, k)
    PRELOAD(k_reset);
    PRELOAD(k_raygen);
    PRELOAD((k_trace_persistent<false, NoCount, FLX_TRACE_BLOCK, false, 9, 0>));
    PRELOAD((k_trace_persistent<true, NoCount, FLX_TRACE_BLOCK, false, 10, 0>));
    PRELOAD((k_logic<false, 3, 2>));
    PRELOAD((k_logic<false, 3, 2, FLX_LOGIC_TILE, FLX_BXDF_DIFFUSE>));
    PRELOAD((k_logic<true, 3, 1>));
    PRELOAD((k_logic<true, 3, 2, FLX_LOGIC_TILE, FLX_CHEAP_BXDF>));
    PRELOAD((k_logic<false, 3, 0>));
    PRELOAD((k_logic<true, 3, 0>));
    constexpr int ALL = FLX_BXDF_DIFFUSE | FLX_BXDF_GLOSSY | FLX_BXDF_GGX_ROUGH_REFLECTION | FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_GGX_ROUGH_DIELECTRIC |
                        FLX_BXDF_IDEAL_DIELECTRIC | FLX_BXDF_EMISSIVE;
    PRELOAD(k_material<ALL>);
    PRELOAD(k_material<FLX_BXDF_DIFFUSE>);
    PRELOAD(k_material<FLX_BXDF_GLOSSY>);
    PRELOAD(k_material<FLX_BXDF_GGX_ROUGH_REFLECTION>);
    PRELOAD(k_material<FLX_BXDF_GGX_ROUGH_DIELECTRIC>);
    PRELOAD((k_material<FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC>));
    PRELOAD(k_postprocess);
    PRELOAD(k_end_iteration);
    PRELOAD(k_repack_scatter_treelet);
    PRELOAD(k_repack_flags);
    PRELOAD(k_repack_emit);
    PRELOAD(k_validate_triangles);
    PRELOAD(k_build_tattr);
#undef PRELOAD
    cudaGetLastError();
}

// ================================================================================================ C ABI
// No C++ exception may cross the C ABI (std::bad_alloc from a host-side staging vector would otherwise end in std::terminate in
// the caller's process): every int-returning entry point is a function-try-block that turns it into an error code + message.
#define FLX_API_CATCH(c)                                                                                               \
    catch (const std::exception &e_)                                                                                   \
    {                                                                                                                  \
        return fail((c), FLX_E_INVALID, "out of memory or internal error: %s", e_.what());                             \
    }                                                                                                                  \
    catch (...)                                                                                                        \
    {                                                                                                                  \
        return fail((c), FLX_E_INVALID, "internal error");                                                             \
    }

extern "C"
{
const char *flx_version(void) { return "fluctus_b200 0.1 (sm_100a)"; }

const char *flx_last_error(const flx_ctx *ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int flx_create(int device, uint32_t num_tasks, flx_ctx **out)
try
{
    flx_ctx *ctx = nullptr; // for the CU/REQUIRE macros: errors land in the thread-local create message
    if (!out || num_tasks == 0)
        return fail(nullptr, FLX_E_INVALID, "flx_create: bad arguments");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, FLX_E_NO_DEVICE, "no CUDA device: %s (this library has no CPU path)", cudaGetErrorString(e));
    if (device < 0 || device >= count)
        return fail(nullptr, FLX_E_INVALID, "device %d out of range (%d devices)", device, count);
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(nullptr, FLX_E_UNSUPPORTED_ARCH, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CU(cudaSetDevice(device));
    flx_ctx *c = new flx_ctx();
    c->device = device;
    c->numTasks = num_tasks;
    memset(&c->params, 0, sizeof c->params);
    ctx = c;
    auto bail = [&](int code) {
        g_create_error = c->error;
        flx_destroy(c);
        return code;
    };
#define CUB(call)                                                                                                      \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
        {                                                                                                              \
            fail(c, (int)e_, "%s failed: %s", #call, cudaGetErrorString(e_));                                          \
            return bail((int)e_);                                                                                      \
        }                                                                                                              \
    } while (0)
    // The main stream gets the highest priority, the two side streams the lowest: when the shadow-ray kernel (stream2) and the
    // extension kernel (main) are both ready, the block scheduler places the extension kernel's CTAs first -- it is the longer of
    // the two, and the shadow kernel then fills the SMs its tail frees.  Without priorities the order is a coin toss (and with
    // the shadow kernel first the extension kernel's own elapsed time, the roofline's denominator, includes its wait).
    int prioLow = 0, prioHigh = 0;
    CUB(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
    CUB(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prioHigh));
    CUB(cudaStreamCreateWithPriority(&c->stream2, cudaStreamNonBlocking, prioLow));
    CUB(cudaStreamCreateWithPriority(&c->stream3, cudaStreamNonBlocking, prioLow));
    CUB(cudaEventCreateWithFlags(&c->evPostFork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->evPostJoin, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->evPixels, cudaEventDisableTiming));
    c->cur = c->stream;
    CUB(cudaEventCreateWithFlags(&c->evFork, cudaEventDisableTiming));
    CUB(cudaEventCreateWithFlags(&c->evJoin, cudaEventDisableTiming));
    const size_t taskBytes = (size_t)num_tasks * FLX_NUM_SLOTS * sizeof(uint32_t); // initMCBuffers, clcontext.cpp:116-141
    CUB(cudaMalloc(&c->tasks, taskBytes));
    CUB(cudaMemset(c->tasks, 0, taskBytes));
    for (int i = 0; i < 8; i++)
    {
        CUB(cudaMalloc(&c->queues[i], (size_t)num_tasks * sizeof(uint32_t)));
        CUB(cudaMemset(c->queues[i], 0, (size_t)num_tasks * sizeof(uint32_t)));
    }
    CUB(cudaMalloc(&c->counters, sizeof(flx_QueueCounters)));
    CUB(cudaMemset(c->counters, 0, sizeof(flx_QueueCounters)));
    CUB(cudaMalloc(&c->snapshot, sizeof(flx_QueueCounters)));
    CUB(cudaMemset(c->snapshot, 0, sizeof(flx_QueueCounters)));
    CUB(cudaMalloc(&c->stats, sizeof(flx_RenderStats64)));
    CUB(cudaMemset(c->stats, 0, sizeof(flx_RenderStats64)));
    CUB(cudaMalloc(&c->currPixelIdx, sizeof(uint32_t)));
    CUB(cudaMemset(c->currPixelIdx, 0, sizeof(uint32_t)));
    c->numScanTiles = (num_tasks + 127) / 128; // enough for the smaller of the two tile sizes of k_logic
    CUB(cudaMalloc(&c->scanTiles, (size_t)c->numScanTiles * sizeof(unsigned long long)));
    CUB(cudaMalloc(&c->scanTicket, sizeof(uint32_t)));
    CUB(cudaMallocHost(&c->pinnedCounters, sizeof(flx_QueueCounters) * flx_ctx::kCounterRing));
    CUB(cudaMallocHost(&c->pinnedPixelIdx, sizeof(uint32_t) * flx_ctx::kCounterRing));
    c->numSMs = prop.multiProcessorCount;
    c->maxDynSmem = (int)prop.sharedMemPerBlockOptin;
    CUB(cudaMalloc(&c->fetchCounters, 4 * sizeof(uint32_t)));
    CUB(cudaMalloc(&c->traceCounts, 10 * sizeof(unsigned long long)));
    CUB(cudaMemset(c->traceCounts, 0, 10 * sizeof(unsigned long long)));
    CUB(cudaEventCreate(&c->evStart));
    CUB(cudaEventCreate(&c->evStop));
    // dummy 1x1 environment map (CLContext::setupScene, clcontext.cpp:513-519)
    CUB(cudaMalloc(&c->envRGBA, 4 * sizeof(float)));
    CUB(cudaMemset(c->envRGBA, 0, 4 * sizeof(float)));
    CUB(cudaMalloc(&c->probTable, sizeof(float)));
    CUB(cudaMalloc(&c->pdfTable, sizeof(float)));
    CUB(cudaMalloc(&c->aliasTable, sizeof(int32_t)));
    CUB(cudaMemset(c->probTable, 0, sizeof(float)));
    CUB(cudaMemset(c->pdfTable, 0, sizeof(float)));
    CUB(cudaMemset(c->aliasTable, 0, sizeof(int32_t)));
    preloadKernels();
    CUB(cudaStreamSynchronize(c->stream));
#undef CUB
    *out = c;
    return 0;
}
FLX_API_CATCH((flx_ctx *)nullptr)

void flx_destroy(flx_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    c->pendingStages = 0; // deferred stages nobody asked the result of
    if (c->stream)
        cudaStreamSynchronize(c->stream);
    if (c->gatherStream)
        cudaStreamSynchronize(c->gatherStream);
    flx_comm_destroy(c);
    drainEvents(c);
    for (auto &e : c->freeEvents)
    {
        cudaEventDestroy(e.a);
        cudaEventDestroy(e.b);
    }
    freeDev(c->tasks);
    for (int i = 0; i < 8; i++)
        freeDev(c->queues[i]);
    freeDev(c->counters);
    freeDev(c->snapshot);
    freeDev(c->stats);
    freeDev(c->currPixelIdx);
    freeDev(c->scanTiles);
    freeDev(c->scanTicket);
    freeDev(c->traceCounts);
    freeDev(c->fetchCounters);
    freeDev(c->mkScratch);
    freeDev(c->mkRayQueue);
    freeDev(c->mkRayCount);
    freeDev(c->mkTypeQueues);
    freeDev(c->mkTypeCounts);
    if (c->evStart)
        cudaEventDestroy(c->evStart);
    if (c->evStop)
        cudaEventDestroy(c->evStop);
    if (c->pinnedCounters)
        cudaFreeHost(c->pinnedCounters);
    if (c->pinnedPixelIdx)
        cudaFreeHost(c->pinnedPixelIdx);
    freeDev(c->tris);
    freeDev(c->materials);
    freeDev(c->kdGamma);
    freeDev(c->texDesc);
    freeDev(c->texData);
    freeDev(c->tnodes);
    freeDev(c->ttris);
    freeDev(c->repackPool);
    freeDev(c->tattr);
    freeDev(c->envRGBA);
    freeDev(c->probTable);
    freeDev(c->pdfTable);
    freeDev(c->aliasTable);
    freeDev(c->imageBlock);
    freeDev(c->aovOut);
    freeDev(c->gatherBuf);
    freeDev(c->fullImage);
    freeDev(c->fullPreview);
    freeDev(c->gatherSnapshot);
    if (c->evSnapshot)
        cudaEventDestroy(c->evSnapshot);
    if (c->evGatherDone)
        cudaEventDestroy(c->evGatherDone);
    for (cudaEvent_t e : c->evSnapshotFree)
        if (e)
            cudaEventDestroy(e);
    if (c->gatherStream)
        cudaStreamDestroy(c->gatherStream);
    if (c->evFork)
        cudaEventDestroy(c->evFork);
    if (c->evJoin)
        cudaEventDestroy(c->evJoin);
    if (c->evPostFork)
        cudaEventDestroy(c->evPostFork);
    if (c->evPostJoin)
        cudaEventDestroy(c->evPostJoin);
    if (c->evPixels)
        cudaEventDestroy(c->evPixels);
    if (c->stream3)
        cudaStreamDestroy(c->stream3);
    if (c->stream2)
        cudaStreamDestroy(c->stream2);
    if (c->stream)
        cudaStreamDestroy(c->stream);
    delete c;
}

uint32_t flx_num_tasks(const flx_ctx *ctx) { return ctx ? ctx->numTasks : 0; }
uint32_t flx_tile_pixels(const flx_ctx *ctx) { return ctx ? ctx->tilePixels : 0; }

// Page-locked host memory for the arrays a caller hands to flx_upload_scene / receives from flx_read_pixels: those copies then
// go by DMA at link speed instead of through the driver's staging buffer (the reference's analogue: the GL pixel-buffer object
// its kernels write the picture into, clcontext.cpp:326-384 -- no pageable host copy there either).
int flx_host_alloc(void **out, size_t bytes)
try
{
    if (!out || bytes == 0)
        return fail(nullptr, FLX_E_INVALID, "flx_host_alloc: bad arguments");
    *out = nullptr;
    const cudaError_t e = cudaHostAlloc(out, bytes, cudaHostAllocPortable);
    if (e != cudaSuccess)
        return fail(nullptr, (int)e, "flx_host_alloc: cudaHostAlloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    return 0;
}
FLX_API_CATCH((flx_ctx *)nullptr)

void flx_host_free(void *p)
{
    if (p)
        cudaFreeHost(p);
}

size_t flx_device_bytes(const flx_ctx *ctx)
{
    if (!ctx)
        return 0;
    return (size_t)ctx->numTasks * (FLX_NUM_SLOTS + 8) * 4 + ctx->sceneBytes + (size_t)ctx->tilePixels * 48;
}

static int uploadSceneImpl(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, const uint32_t *indices, uint32_t n_indices, const flx_Node *nodes,
                           uint32_t n_nodes, const flx_Material *materials, uint32_t n_materials, const flx_TexDescriptor *tex_desc, uint32_t n_tex,
                           const uint8_t *tex_data, size_t tex_bytes);

int flx_upload_scene(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, const uint32_t *indices, uint32_t n_indices, const flx_Node *nodes,
                     uint32_t n_nodes, const flx_Material *materials, uint32_t n_materials, const flx_TexDescriptor *tex_desc, uint32_t n_tex,
                     const uint8_t *tex_data, size_t tex_bytes)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    const int rc = uploadSceneImpl(ctx, tris, n_tris, indices, n_indices, nodes, n_nodes, materials, n_materials, tex_desc, n_tex, tex_data, tex_bytes);
    if (rc) // the copies are asynchronous: whatever was enqueued before the failure must be done with the caller's arrays
        cudaStreamSynchronize(ctx->stream);
    return rc;
}
FLX_API_CATCH(ctx)

static int uploadSceneImpl(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, const uint32_t *indices, uint32_t n_indices, const flx_Node *nodes,
                           uint32_t n_nodes, const flx_Material *materials, uint32_t n_materials, const flx_TexDescriptor *tex_desc, uint32_t n_tex,
                           const uint8_t *tex_data, size_t tex_bytes)
{
    REQUIRE(tris && indices && nodes && materials, "flx_upload_scene: null array");
    REQUIRE(n_tris > 0 && n_indices > 0 && n_nodes > 0 && n_materials > 0, "flx_upload_scene: empty scene");
    REQUIRE(n_tex == 0 || (tex_desc && tex_data), "flx_upload_scene: texture descriptors without data");
    if (ctx->repackOnHost) // the device repack checks the material indices on the device copy (k_validate_triangles)
        for (uint32_t i = 0; i < n_tris; i++)
            if (tris[i].matId < 0 || (uint32_t)tris[i].matId >= n_materials)
                return fail(ctx, FLX_E_INVALID, "triangle %u has a material index outside the %u uploaded", i, n_materials);
    for (uint32_t i = 0; i < n_materials; i++)
    {
        const int maps[3] = {materials[i].map_Kd, materials[i].map_Ks, materials[i].map_N};
        for (int m : maps)
            if (m < -1 || m >= (int)n_tex)
                return fail(ctx, FLX_E_INVALID, "material %u references texture %d of %u", i, m, n_tex);
    }
    for (uint32_t i = 0; i < n_tex; i++)
        if ((size_t)tex_desc[i].offset + (size_t)tex_desc[i].width * tex_desc[i].height * 4 > tex_bytes || (tex_desc[i].offset & 3u) || tex_desc[i].width == 0 ||
            tex_desc[i].height == 0)
            return fail(ctx, FLX_E_INVALID, "texture %u: descriptor outside the %zu-byte blob", i, tex_bytes);
    {
        // Cheap fingerprint for flx_checkpoint_save / _load: the sizes, every material, and a sample of nodes and triangles
        // (hashing the ~60 MB of a scene in full would cost more than uploading it).
        uint64_t hsh = 14695981039346656037ull;
        const uint32_t sizes[5] = {n_tris, n_indices, n_nodes, n_materials, n_tex};
        hsh = hsh * 1099511628211ull ^ 0;
        for (uint32_t v : sizes)
            hsh = (hsh ^ v) * 1099511628211ull;
        for (uint32_t i = 0; i < n_materials; i++)
            for (size_t b = 0; b < 72; b++) // up to and including `type`; the struct's tail is padding
                hsh = (hsh ^ reinterpret_cast<const unsigned char *>(materials + i)[b]) * 1099511628211ull;
        for (uint32_t i = 0; i < n_nodes; i += std::max(1u, n_nodes / 512u))
            for (size_t b = 0; b < 41; b++)
                hsh = (hsh ^ reinterpret_cast<const unsigned char *>(nodes + i)[b]) * 1099511628211ull;
        for (uint32_t i = 0; i < n_tris; i += std::max(1u, n_tris / 512u))
            for (size_t b = 0; b < 12; b++)
                hsh = (hsh ^ reinterpret_cast<const unsigned char *>(&tris[i].v0.p)[b]) * 1099511628211ull;
        ctx->sceneHash = hsh;
    }
    const bool timing = std::getenv("FLX_DEBUG_TIMING") != nullptr; // prints where the upload's wall time goes
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
    const auto t0 = now();
    Repacked rp;
    int rc = 0;
    if (ctx->repackOnHost && (rc = repackBvh(ctx, tris, n_tris, indices, n_indices, nodes, n_nodes, rp)))
        return rc;
    const auto t1 = now();
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->sceneReady = false;
    ctx->sceneBytes = 0;
    auto lap = [&](const char *what) { // FLX_DEBUG_TIMING: where the upload's time goes (synchronises, so only when asked for)
        static std::chrono::steady_clock::time_point last;
        if (!timing)
            return;
        cudaStreamSynchronize(ctx->stream);
        const auto t = now();
        if (what)
            std::fprintf(stderr, "  flx_upload_scene: %-28s %8.3f ms\n", what, ms(last, t));
        last = t;
    };
    lap(nullptr);
    if ((rc = uploadArray(ctx, ctx->tris, tris, n_tris)))
        return rc;
    if ((rc = reserveDev(ctx, ctx->tattr, (size_t)n_tris * 64)))
        return rc;
    ctx->sceneBytes += (size_t)n_tris * 64;
    k_build_tattr<<<(n_tris + 255) / 256, 256, 0, ctx->stream>>>(ctx->tris, n_tris, ctx->tattr);
    if ((rc = launchCheck(ctx, "k_build_tattr")))
        return rc;
    lap("triangles (malloc + copy)");
    if ((rc = uploadArray(ctx, ctx->materials, materials, n_materials)))
        return rc;
    ctx->materialTypes = 0;
    ctx->otherTypes = false;
    for (uint32_t i = 0; i < n_materials; i++)
    {
        const int ty = materials[i].type;
        ctx->materialTypes |= (uint32_t)ty;
        if (ty != FLX_BXDF_DIFFUSE && ty != FLX_BXDF_GLOSSY && ty != FLX_BXDF_GGX_ROUGH_REFLECTION && ty != FLX_BXDF_GGX_ROUGH_DIELECTRIC && ty != FLX_BXDF_IDEAL_REFLECTION &&
            ty != FLX_BXDF_IDEAL_DIELECTRIC)
            ctx->otherTypes = true;
    }
    std::vector<float4> kdGamma(n_materials);
    for (uint32_t i = 0; i < n_materials; i++) // matGetAlbedo of an untextured material (utils.cl:136-141), hoisted out of the kernels
        kdGamma[i] = make_float4(flx_powf(materials[i].Kd.x, 2.2f), flx_powf(materials[i].Kd.y, 2.2f), flx_powf(materials[i].Kd.z, 2.2f), 0.0f);
    if ((rc = uploadArray(ctx, ctx->kdGamma, kdGamma.data(), n_materials)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->texDesc, tex_desc, n_tex)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->texData, tex_data, tex_bytes, 4)))
        return rc;
    ctx->nTris = n_tris;
    lap("materials + textures");
    if (!ctx->repackOnHost)
    {
        if ((rc = repackOnDevice(ctx, indices, n_indices, nodes, n_nodes, n_tris, n_materials)))
            return rc;
        lap("hierarchy copy + repack");
        ctx->sceneReady = true;
        if (ctx->l2Persist && (rc = applyL2Persist(ctx)))
            return rc;
        if (timing)
            std::fprintf(stderr, "flx_upload_scene: allocations + copies + repack on the device %.2f ms (%.1f MB resident)\n", ms(t1, now()), ctx->sceneBytes / 1e6);
        return 0;
    }
    if ((rc = uploadArray(ctx, ctx->tnodes, rp.nodes.data(), rp.nodes.size(), 4)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->ttris, rp.tris.data(), rp.tris.size(), 4)))
        return rc;
    ctx->rootRef = rp.rootRef;
    ctx->nTNodes = (uint32_t)(rp.nodes.size() / 4);
    ctx->nTTris = (uint32_t)(rp.tris.size() / 4);
    ctx->treeletNodes = rp.treeletNodes;
    CU(cudaStreamSynchronize(ctx->stream)); // the copies above borrow the caller's (and this function's) host arrays
    ctx->sceneReady = true;
    if (ctx->l2Persist && (rc = applyL2Persist(ctx)))
        return rc;
    if (timing)
        std::fprintf(stderr, "flx_upload_scene: repack on the host %.2f ms, allocations + copies %.2f ms (%.1f MB)\n", ms(t0, t1), ms(t1, now()), ctx->sceneBytes / 1e6);
    return 0;
}

// ---- GPU hierarchy builder (flx_bvh_build.cuh): stands in for `new SBVH(&tris, mode)` / BVH::m_nodes + m_indices
// (src/scene.cpp:574-590, src/sbvh.cpp:4-73) when build time matters more than tree quality
int flx_build_bvh(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, uint32_t max_leaf, int quality, flx_Node *nodes_out, uint32_t nodes_capacity,
                  uint32_t *n_nodes_out, uint32_t *indices_out, float *build_ms)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(tris && nodes_out && n_nodes_out && indices_out, "flx_build_bvh: null array");
    REQUIRE(n_tris > 0 && n_tris < 0x40000000u, "flx_build_bvh: triangle count out of range");
    REQUIRE(max_leaf >= 1 && max_leaf <= 255, "flx_build_bvh: max_leaf must be in 1..255 (nPrims is a byte, src/bvhnode.hpp:58)");
    REQUIRE(quality == FLX_BVH_FAST || quality == FLX_BVH_PLOC || quality == FLX_BVH_PLOC_OPT, "flx_build_bvh: quality must be FLX_BVH_FAST, FLX_BVH_PLOC or FLX_BVH_PLOC_OPT");
    const bool ploc = quality != FLX_BVH_FAST;
    const int reinsertIterations = quality == FLX_BVH_PLOC_OPT && n_tris > 2 ? ctx->bvhReinsertIterations : 0;
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = n_tris, total = 2u * n - 1u;
    BvhBuild b;
    memset(&b, 0, sizeof b);
    b.n = n;
    b.maxLeaf = max_leaf;
    b.triCost = (float)ctx->bvhTriCostPercent / 100.0f;
    flx_Triangle *dTris = nullptr;
    void *sortTemp = nullptr;
    size_t sortBytes = 0;
    std::vector<void *> owned;
    auto release = [&]() {
        for (void *p : owned)
            cudaFree(p);
    };
#define BALLOC(ptr, count)                                                                                                                                      \
    do                                                                                                                                                         \
    {                                                                                                                                                          \
        void *p_ = nullptr;                                                                                                                                    \
        cudaError_t e_ = cudaMalloc(&p_, std::max<size_t>((size_t)(count) * sizeof(*(ptr)), 16));                                                              \
        if (e_ != cudaSuccess)                                                                                                                                 \
        {                                                                                                                                                      \
            release();                                                                                                                                         \
            return fail(ctx, (int)e_, "flx_build_bvh: cudaMalloc failed: %s", cudaGetErrorString(e_));                                                         \
        }                                                                                                                                                      \
        owned.push_back(p_);                                                                                                                                   \
        (ptr) = reinterpret_cast<decltype(ptr)>(p_);                                                                                                           \
    } while (0)
    BALLOC(dTris, n);
    BALLOC(b.keys, n);
    BALLOC(b.keysSorted, n);
    BALLOC(b.bmin, total);
    BALLOC(b.bmax, total);
    BALLOC(b.primMin, n);
    BALLOC(b.primMax, n);
    BALLOC(b.parent, total);
    BALLOC(b.children, n);
    BALLOC(b.range, n);
    BALLOC(b.cost, total);
    BALLOC(b.size, total);
    BALLOC(b.collapsed, total); // indexed by node id in k_bvh_emit; leaves are never collapsed
    BALLOC(b.visits, n);
    BALLOC(b.sceneBounds, 6);
    BALLOC(b.nodesOut, total);
    BALLOC(b.indicesOut, n);
    b.tris = dTris;
    PlocBuild pb;
    memset(&pb, 0, sizeof pb);
    void *scanTemp = nullptr;
    size_t scanBytes = 0;
    ReinsertView ri;
    memset(&ri, 0, sizeof ri);
    if (ploc)
    {
        pb.n = n;
        pb.maxLeaf = max_leaf;
        pb.triCost = b.triCost;
        pb.keysSorted = b.keysSorted;
        pb.primMin = b.primMin;
        pb.primMax = b.primMax;
        pb.bmin = b.bmin;
        pb.bmax = b.bmax;
        pb.parent = b.parent;
        pb.cost = b.cost;
        pb.size = b.size;
        pb.collapsed = b.collapsed;
        pb.nodesOut = b.nodesOut;
        pb.indicesOut = b.indicesOut;
        BALLOC(pb.left, total);
        BALLOC(pb.right, total);
        BALLOC(pb.prims, total);
        BALLOC(pb.cidA, n);
        BALLOC(pb.cidB, n);
        BALLOC(pb.nn, n);
        BALLOC(pb.flags, n);
        BALLOC(pb.scan, n);
        BALLOC(pb.depthMax, 1);
        BALLOC(pb.state, 2);
        if (reinsertIterations > 0)
        {
            BALLOC(ri.alive, total);
            BALLOC(ri.gain, total);
            BALLOC(ri.out, total);
            BALLOC(ri.pivot, total);
            BALLOC(ri.lock, total);
            BALLOC(ri.cand, total);
            BALLOC(ri.win, total);
            BALLOC(ri.visits, total);
            BALLOC(ri.moves, 1);
        }
        cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, pb.flags, pb.scan, (int)n, ctx->stream);
        unsigned char *t = nullptr;
        BALLOC(t, scanBytes);
        scanTemp = t;
    }
    cub::DeviceRadixSort::SortKeys(nullptr, sortBytes, b.keys, b.keysSorted, (int)n, 0, 62, ctx->stream);
    {
        unsigned char *t = nullptr;
        BALLOC(t, sortBytes);
        sortTemp = t;
    }
#undef BALLOC
    cudaStream_t st = ctx->stream;
    int rc = 0;
    auto cu = [&](cudaError_t e, const char *what) {
        if (e != cudaSuccess && rc == 0)
            rc = fail(ctx, (int)e, "flx_build_bvh: %s failed: %s", what, cudaGetErrorString(e));
    };
    cu(cudaMemcpyAsync(dTris, tris, (size_t)n * sizeof(flx_Triangle), cudaMemcpyHostToDevice, st), "triangle upload");
    cu(cudaStreamSynchronize(st), "triangle upload");
    cu(cudaEventRecord(ctx->evStart, st), "event");
    const uint32_t boundsInit[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    cu(cudaMemcpyAsync(b.sceneBounds, boundsInit, sizeof boundsInit, cudaMemcpyHostToDevice, st), "bounds init");
    cu(cudaMemsetAsync(b.parent, 0xff, (size_t)total * sizeof(int), st), "memset");
    cu(cudaMemsetAsync(b.visits, 0, (size_t)n * sizeof(uint32_t), st), "memset");
    cu(cudaMemsetAsync(b.collapsed, 0, (size_t)total * sizeof(uint32_t), st), "memset");
    const unsigned gridN = (n + FLX_BVH_BLOCK - 1) / FLX_BVH_BLOCK, gridT = (total + FLX_BVH_BLOCK - 1) / FLX_BVH_BLOCK;
    if (rc == 0)
    {
        k_bvh_prims<<<gridN, FLX_BVH_BLOCK, 0, st>>>(b);
        k_bvh_morton<<<gridN, FLX_BVH_BLOCK, 0, st>>>(b);
        cu(cub::DeviceRadixSort::SortKeys(sortTemp, sortBytes, b.keys, b.keysSorted, (int)n, 0, 62, st), "radix sort");
        if (ploc)
        {
            cu(cudaMemsetAsync(pb.depthMax, 0, sizeof(uint32_t), st), "memset");
            k_ploc_init<<<gridN, FLX_BVH_BLOCK, 0, st>>>(pb);
            // Rounds are enqueued in batches without waiting: the cluster count and the next node id live on the device, every
            // kernel reads them there.  After a batch the host fetches the count -- to stop, and to shrink the grids.  A round
            // with one cluster left is a no-op, so overshooting inside a batch is harmless.
            uint32_t bound = n; // cluster count as last seen by the host
            uint32_t *cid = pb.cidA, *cidNext = pb.cidB;
            const int kBatch = 8;
            while (bound > 1 && rc == 0)
            {
                const unsigned gridM = (bound + FLX_BVH_BLOCK - 1) / FLX_BVH_BLOCK;
                for (int r = 0; r < kBatch; r++)
                {
                    k_ploc_nearest<<<gridM, FLX_BVH_BLOCK, 0, st>>>(pb, cid);
                    k_ploc_flags<<<gridM, FLX_BVH_BLOCK, 0, st>>>(pb, bound);
                    cu(cub::DeviceScan::ExclusiveSum(scanTemp, scanBytes, pb.flags, pb.scan, (int)bound, st), "scan");
                    k_ploc_apply<<<gridM, FLX_BVH_BLOCK, 0, st>>>(pb, cid, cidNext);
                    k_ploc_advance<<<1, 32, 0, st>>>(pb);
                    std::swap(cid, cidNext);
                }
                uint32_t left = 0;
                cu(cudaMemcpyAsync(&left, pb.state, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "round read-back");
                cu(cudaStreamSynchronize(st), "PLOC rounds");
                if (rc == 0 && (left == 0 || left >= bound)) // every round merges at least the closest pair
                    rc = fail(ctx, FLX_E_INVALID, "flx_build_bvh: PLOC made no progress (%u clusters after a batch that started with %u)", left, bound);
                bound = left;
            }
            if (reinsertIterations > 0 && rc == 0)
            {
                // the post-pass of FLX_BVH_PLOC_OPT (flx_bvh_build.cuh, "parallel reinsertion"): nothing comes back to the host in between
                ri.b = pb;
                ri.root = (int)(total - 1u); // the last node created
                ri.total = total;
                cu(cudaMemsetAsync(ri.moves, 0, sizeof(uint32_t), st), "memset");
                k_ri_alive<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                for (int it = 0; it < reinsertIterations; it++)
                {
                    k_ri_search<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    k_ri_lock<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    k_ri_cand<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    k_ri_guard<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    k_ri_apply<<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    if (it + 1 < reinsertIterations)
                        k_ri_refit<false><<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                    else
                        k_ri_refit<true><<<gridT, FLX_BVH_BLOCK, 0, st>>>(ri);
                }
            }
            k_ploc_emit<<<gridT, FLX_BVH_BLOCK, 0, st>>>(pb);
        }
        else
        {
            if (n > 1)
                k_bvh_hierarchy<<<gridN, FLX_BVH_BLOCK, 0, st>>>(b);
            k_bvh_fit<<<gridN, FLX_BVH_BLOCK, 0, st>>>(b);
            k_bvh_emit<<<gridT, FLX_BVH_BLOCK, 0, st>>>(b);
        }
        cu(cudaGetLastError(), "kernel launch");
        cu(cudaEventRecord(ctx->evStop, st), "event");
    }
    uint32_t nNodes = 0;
    if (rc == 0)
    {
        // the root: LBVH internal node 0 (or the single leaf, also id 0); PLOC: the last node created
        const uint32_t rootId = ploc ? total - 1u : 0u;
        uint32_t depth = 0;
        cu(cudaMemcpyAsync(&nNodes, b.size + rootId, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "size read-back");
        if (ploc)
            cu(cudaMemcpyAsync(&depth, pb.depthMax, sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "depth read-back");
        cu(cudaStreamSynchronize(st), "build");
        // the traversal stack holds 64 entries (src/bvh.cl:240); a radix tree cannot get there, these trees could
        if (rc == 0 && depth > (uint32_t)ctx->bvhDepthLimit && reinsertIterations > 0)
        {
            // reinsertion may deepen a tree (Country Kitchen: 55 -> 56 levels): fall back to the tree it started from
            release();
            return flx_build_bvh(ctx, tris, n_tris, max_leaf, FLX_BVH_PLOC, nodes_out, nodes_capacity, n_nodes_out, indices_out, build_ms);
        }
        if (rc == 0 && depth > (uint32_t)ctx->bvhDepthLimit)
            rc = fail(ctx, FLX_E_INVALID, "flx_build_bvh: PLOC tree is %u levels deep (limit %d); use FLX_BVH_FAST for this input", depth, ctx->bvhDepthLimit);
    }
    if (rc == 0 && nNodes > nodes_capacity)
        rc = fail(ctx, FLX_E_INVALID, "flx_build_bvh: %u nodes do not fit the caller's %u", nNodes, nodes_capacity);
    if (rc == 0)
    {
        cu(cudaMemcpyAsync(nodes_out, b.nodesOut, (size_t)nNodes * sizeof(flx_Node), cudaMemcpyDeviceToHost, st), "node read-back");
        cu(cudaMemcpyAsync(indices_out, b.indicesOut, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st), "index read-back");
        cu(cudaStreamSynchronize(st), "read-back");
        *n_nodes_out = nNodes;
        if (build_ms)
        {
            float ms = 0.0f;
            cu(cudaEventElapsedTime(&ms, ctx->evStart, ctx->evStop), "event");
            *build_ms = ms;
        }
    }
    release();
    return rc;
}
FLX_API_CATCH(ctx)

int flx_upload_envmap(flx_ctx *ctx, const float *rgb, int32_t w, int32_t h, const float *prob, const int32_t *alias, const float *pdf)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(rgb && prob && alias && pdf && w > 0 && h > 0, "flx_upload_envmap: bad arguments");
    const size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; i++)
        if (alias[i] < 0 || (size_t)alias[i] >= n)
            return fail(ctx, FLX_E_INVALID, "alias table entry %zu = %d out of range", i, alias[i]);
    std::vector<float> rgba(n * 4); // RGB -> RGBA with alpha 1 (clcontext.cpp:472-487)
    for (size_t i = 0; i < n; i++)
    {
        rgba[4 * i + 0] = rgb[3 * i + 0];
        rgba[4 * i + 1] = rgb[3 * i + 1];
        rgba[4 * i + 2] = rgb[3 * i + 2];
        rgba[4 * i + 3] = 1.0f;
    }
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t before = ctx->sceneBytes;
    int rc;
    if ((rc = uploadArray(ctx, ctx->envRGBA, rgba.data(), n * 4)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->probTable, prob, n)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->aliasTable, alias, n)))
        return rc;
    if ((rc = uploadArray(ctx, ctx->pdfTable, pdf, n)))
        return rc;
    (void)before;
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->envW = w;
    ctx->envH = h;
    return 0;
}
FLX_API_CATCH(ctx)

static int allocImage(flx_ctx *ctx)
{
    ctx->pixelsEventValid = false;
    const uint32_t rows = localRows(ctx->height, ctx->part, ctx->nParts, ctx->stripeRows);
    ctx->tilePixels = rows * ctx->width;
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gatherStream)
        CU(cudaStreamSynchronize(ctx->gatherStream));
    ctx->gatherInFlight = false;
    freeDev(ctx->aovOut);
    ctx->pixels = ctx->denoiserAlbedo = ctx->denoiserNormal = ctx->preview = nullptr;
    ctx->dirtyPixels = nullptr;
    // the gather buffers (flx_gather_pixels) are capacity-checked there, each on its own: they stay
    ctx->snapshotBusy[0] = ctx->snapshotBusy[1] = false;
    ctx->previewStale = true;
    if (ctx->tilePixels == 0)
        return fail(ctx, FLX_E_INVALID, "tile %u of %u owns no rows of a %ux%u image", ctx->part, ctx->nParts, ctx->width, ctx->height);
    // one allocation (kept while large enough), two asynchronous fills (the reference makes three buffers + two GL PBOs, clcontext.cpp:326-384)
    const size_t bytes = (size_t)ctx->tilePixels * 4 * sizeof(float);
    if (int rcImage = reserveDev(ctx, ctx->imageBlock, 4 * bytes + ctx->tilePixels))
        return rcImage;
    ctx->pixels = reinterpret_cast<float *>(ctx->imageBlock);
    ctx->denoiserAlbedo = reinterpret_cast<float *>(ctx->imageBlock + bytes);
    ctx->denoiserNormal = reinterpret_cast<float *>(ctx->imageBlock + 2 * bytes);
    ctx->preview = reinterpret_cast<float *>(ctx->imageBlock + 3 * bytes);
    ctx->dirtyPixels = ctx->imageBlock + 4 * bytes;
    CU(cudaMemsetAsync(ctx->imageBlock, 0, 4 * bytes, ctx->stream));
    CU(cudaMemsetAsync(ctx->dirtyPixels, 1, ctx->tilePixels, ctx->stream));
    return 0;
}

int flx_resize(flx_ctx *ctx, uint32_t width, uint32_t height)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(width > 0 && height > 0 && (uint64_t)width * height < 0x7fffffffull, "flx_resize: bad image size");
    ctx->width = width;
    ctx->height = height;
    return allocImage(ctx);
}
FLX_API_CATCH(ctx)

int flx_set_tile(flx_ctx *ctx, uint32_t part, uint32_t n_parts, uint32_t stripe_rows)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(n_parts > 0 && part < n_parts && stripe_rows > 0, "flx_set_tile: bad arguments");
    ctx->part = part;
    ctx->nParts = n_parts;
    ctx->stripeRows = stripe_rows;
    if (ctx->width && ctx->height)
        return allocImage(ctx);
    return 0;
}
FLX_API_CATCH(ctx)

int flx_update_params(flx_ctx *ctx, const flx_RenderParams *p)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(p != nullptr, "flx_update_params: null params");
    REQUIRE(p->width > 0 && p->height > 0, "flx_update_params: empty image");
    if (ctx->pixels && (p->width != ctx->width || p->height != ctx->height))
        return fail(ctx, FLX_E_INVALID, "params are %ux%u but pixel storage is %ux%u: call flx_resize first", p->width, p->height, ctx->width, ctx->height);
    if (!ctx->paramsSet || p->ppParams.exposure != ctx->params.ppParams.exposure || p->ppParams.tmOperator != ctx->params.ppParams.tmOperator)
        ctx->previewStale = true; // the display pass maps every pixel differently now
    ctx->params = *p;
    ctx->tanHalfFov = flx_tanf(0.5f * p->camera.fov * 3.14159265358979323846f / 180); // toRad, geom.h:22; wf_raygen.cl:50
    {
        // lens_offset (flx_kernels.cuh): scale * (right * rx + up * ry) is +-0 for every disk sample iff the scale is 0 and the basis
        // is finite; adding +-0 leaves the origin's bits alone iff no component of it is -0.0f
        auto finite3 = [](const flx_float3 &v) { return std::isfinite(v.x) && std::isfinite(v.y) && std::isfinite(v.z); };
        auto negZero = [](float x) { return x == 0.0f && std::signbit(x); };
        const flx_Camera &c = p->camera;
        ctx->pinholeCamera = (p->worldRadius * c.apertureSize == 0.0f) && std::isfinite(p->worldRadius) && finite3(c.right) && finite3(c.up) && !negZero(c.pos.x) &&
                             !negZero(c.pos.y) && !negZero(c.pos.z);
    }
    ctx->paramsSet = true;
    return 0;
}
FLX_API_CATCH(ctx)

int flx_enqueue_reset(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = std::max(ctx->numTasks, ctx->tilePixels); // clcontext.cpp:767
    Timed tm(ctx, FLX_K_RESET);
    k_reset<<<(n + FLX_BLOCK - 1) / FLX_BLOCK, FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params);
    markPixelsWritten(ctx);
    return launchCheck(ctx, "k_reset");
}
FLX_API_CATCH(ctx)

static int launchRaygen(flx_ctx *ctx)
{
    Timed tm(ctx, FLX_K_RAYGEN);
    k_raygen<<<streamingGrid(ctx->numTasks), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params);
    return launchCheck(ctx, "k_raygen");
}

int flx_enqueue_raygen(flx_ctx *ctx)
try
{
    const bool afterLogic = ctx && ctx->pendingStages == 1;
    int rc = checkReady(ctx, false, true, afterLogic);
    if (rc)
        return rc;
    if (afterLogic) // logic is pending and this is the call that follows it in the reference's loop: keep deferring
    {
        ctx->pendingStages = 2;
        return 0;
    }
    CU(cudaSetDevice(ctx->device));
    return launchRaygen(ctx);
}
FLX_API_CATCH(ctx)

int flx_enqueue_extrays(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    if (ctx->overlapTrace)
        CU(cudaEventRecord(ctx->evFork, ctx->stream)); // a shadow-ray stage enqueued next may start from here (see flx_ctx::opSeq)
    ctx->extSeq = ctx->opSeq;
    if (ctx->traceVariant >= 1)
        return launchPersistent<false>(ctx);
    Timed tm(ctx, FLX_K_EXTRAYS);
    const unsigned grid = (ctx->numTasks + FLX_TRACE_BLOCK - 1) / FLX_TRACE_BLOCK;
    if (ctx->counting)
        k_extrays<RayCount><<<grid, FLX_TRACE_BLOCK, 0, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, ctx->traceCounts);
    else
        k_extrays<NoCount><<<grid, FLX_TRACE_BLOCK, 0, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->tris, nullptr);
    return launchCheck(ctx, "k_extrays");
}
FLX_API_CATCH(ctx)

static int launchShadow(flx_ctx *ctx)
{
    if (ctx->traceVariant >= 1)
        return launchPersistent<true>(ctx);
    Timed tm(ctx, FLX_K_SHADOWRAYS);
    const unsigned grid = (ctx->numTasks + FLX_TRACE_BLOCK - 1) / FLX_TRACE_BLOCK;
    if (ctx->counting)
        k_shadowrays<RayCount><<<grid, FLX_TRACE_BLOCK, 0, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), ctx->traceCounts + 5);
    else
        k_shadowrays<NoCount><<<grid, FLX_TRACE_BLOCK, 0, ctx->cur>>>(makeFrame(ctx), ctx->params, makeBvh(ctx), nullptr);
    return launchCheck(ctx, "k_shadowrays");
}

int flx_enqueue_shadowrays(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    if (!(ctx->overlapTrace && ctx->extSeq + 1 == ctx->opSeq))
        return launchShadow(ctx);
    // directly after the extension stage: run beside it on the second stream, so its CTAs fill the SMs the extension kernel's
    // tail leaves idle; everything enqueued later on the main stream waits for both
    CU(cudaStreamWaitEvent(ctx->stream2, ctx->evFork, 0));
    ctx->cur = ctx->stream2;
    rc = launchShadow(ctx);
    ctx->cur = ctx->stream;
    if (rc)
        return rc;
    CU(cudaEventRecord(ctx->evJoin, ctx->stream2));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->evJoin, 0));
    return 0;
}
FLX_API_CATCH(ctx)

static int launchMaterials(flx_ctx *ctx, uint32_t doneInLogic = 0u);
// wf_logic alone, or (fused) wf_logic + wf_raygen + wf_mat_* in one pass over the path state (k_logic<.., FUSE>)
static int launchLogic(flx_ctx *ctx, int first_iteration, bool fused)
{
    const uint32_t maxId = first_iteration ? std::min(ctx->tilePixels, ctx->numTasks) : ctx->numTasks; // wf_logic.cl:45
    const uint32_t LT = ctx->logicTile == 128 ? 128u : 256u;
    const uint32_t tiles = (maxId + LT - 1) / LT;
    // flx_render: the previous iteration's last kernel has zeroed every status word the previous logic launch touched, which covers this launch
    // unless it uses more tiles (first iteration -> steady state, another tile size)
    if (!ctx->scanClean || tiles > ctx->lastLogicTiles)
    {
        CU(cudaMemsetAsync(ctx->scanTiles, 0, (size_t)std::max(tiles, ctx->lastLogicTiles) * sizeof(unsigned long long), ctx->stream));
        CU(cudaMemsetAsync(ctx->scanTicket, 0, sizeof(uint32_t), ctx->stream));
    }
    ctx->scanClean = false;
    ctx->lastLogicTiles = tiles;
    ScanState scan{ctx->scanTiles, ctx->scanTicket};
    const Frame fr = makeFrame(ctx);
    const SceneView sc = makeScene(ctx);
    const bool sep = ctx->params.wfSeparateQueues != 0;
    bool sepCheap = false;
    if (fused)
    {
        // single material queue: logic + raygen + materials in one kernel; per-type queues: logic + raygen here, then the per-type
        // material kernels (a warp of those sees one BSDF, which is the reason the queues exist)
        {
            Timed tm(ctx, FLX_K_LOGIC_FUSED);
            // the lobes compiled into the material part follow the scene's materials, as the reference's kernel build does
            // (src/kernel_impl.hpp:261-266); FLX_TUNE_MATERIAL_MASK = 0 forces the all-lobes instantiation
            const bool diffuseOnly = ctx->useMaterialMask && ctx->materialTypes == (uint32_t)FLX_BXDF_DIFFUSE;
            // per-type queues exist so that a warp of the material kernels sees ONE heavy BSDF.  When every lobe the scene uses is a cheap one
            // (diffuse, the two ideal ones: Luxball) there is nothing to keep apart: the material part is fused here as well (the per-type queues
            // are still filled, so state, queues and counters are the reference's) and no material kernel runs.  With heavy lobes in the scene
            // as well, fusing just the cheap ones was measured WORSE (Country Kitchen: logic 0.27 -> 0.49 ms for 0.09 ms saved in the material
            // kernels: its diffuse materials are textured and bump-mapped), so that case keeps all its material kernels.
            sepCheap = sep && ctx->useMaterialMask && ctx->materialTypes != 0u && (ctx->materialTypes & ~(uint32_t)FLX_CHEAP_BXDF) == 0u && LT == 256;
#define FUSEDK(MB)                                                                                                                                             \
    do                                                                                                                                                         \
    {                                                                                                                                                          \
        if (sepCheap)                                                                                                                                          \
            k_logic<true, MB, 2, FLX_LOGIC_TILE, FLX_CHEAP_BXDF><<<tiles, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);                      \
        else if (sep)                                                                                                                                          \
            k_logic<true, MB, 1><<<tiles, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);                                                      \
        else if (diffuseOnly)                                                                                                                                  \
            k_logic<false, MB, 2, FLX_LOGIC_TILE, FLX_BXDF_DIFFUSE><<<tiles, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);                   \
        else                                                                                                                                                   \
            k_logic<false, MB, 2><<<tiles, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);                                                     \
    } while (0)
            if (LT == 128) // half-size tiles, compiled for twice as many resident CTAs (same register budget as 3 x 256)
            {
                if (sep)
                    k_logic<true, 6, 1, 128><<<tiles, 128, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);
                else
                    k_logic<false, 6, 2, 128><<<tiles, 128, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);
            }
            else
                switch (ctx->fusedMinBlocks)
                {
                case 4: FUSEDK(4); break;
                case 2: FUSEDK(2); break;
                case 1: FUSEDK(1); break;
                default: FUSEDK(3); break;
                }
#undef FUSEDK
        }
        markPixelsWritten(ctx);
        int rc = launchCheck(ctx, "k_logic<fused>");
        if (rc == 0 && sep)
            rc = launchMaterials(ctx, sepCheap ? (uint32_t)FLX_CHEAP_BXDF : 0u);
        return rc;
    }
    Timed tm(ctx, FLX_K_LOGIC);
#define LOGIC(SEP, MB) k_logic<SEP, MB><<<tiles, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId)
    if (LT == 128)
    {
        if (sep)
            k_logic<true, 6, 0, 128><<<tiles, 128, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);
        else
            k_logic<false, 6, 0, 128><<<tiles, 128, 0, ctx->stream>>>(fr, ctx->params, sc, scan, maxId);
    }
    else
    switch (ctx->logicMinBlocks)
    {
    case 2: if (sep) LOGIC(true, 2); else LOGIC(false, 2); break;
    case 4: if (sep) LOGIC(true, 4); else LOGIC(false, 4); break;
    default: if (sep) LOGIC(true, 3); else LOGIC(false, 3); break;
    }
#undef LOGIC
    markPixelsWritten(ctx);
    return launchCheck(ctx, "k_logic");
}

int flx_enqueue_logic(flx_ctx *ctx, int first_iteration)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    if (ctx->fuseStages) // launched by whatever comes next: fused if that is raygen + materials, on its own otherwise
    {
        ctx->pendingStages = 1;
        ctx->pendingFirstIteration = first_iteration;
        return 0;
    }
    return launchLogic(ctx, first_iteration, false);
}
FLX_API_CATCH(ctx)

extern "C++"
{
namespace
{
int flushPending(flx_ctx *ctx)
{
    const int pending = ctx->pendingStages;
    if (!pending)
        return 0;
    ctx->pendingStages = 0;
    CU(cudaSetDevice(ctx->device));
    int rc = launchLogic(ctx, ctx->pendingFirstIteration, false);
    if (rc == 0 && pending == 2)
        rc = launchRaygen(ctx);
    return rc;
}
} // namespace
} // extern "C++"

int flx_enqueue_materials(flx_ctx *ctx)
try
{
    const bool completesIteration = ctx && ctx->pendingStages == 2;
    int rc = checkReady(ctx, true, true, completesIteration);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    if (completesIteration) // logic, raygen, materials arrived back to back: one pass over the path state does all three
    {
        ctx->pendingStages = 0;
        return launchLogic(ctx, ctx->pendingFirstIteration, true);
    }
    return launchMaterials(ctx);
}
FLX_API_CATCH(ctx)

// doneInLogic: lobes whose paths the fused logic kernel has already taken through their material (their queues are full but need no kernel)
static int launchMaterials(flx_ctx *ctx, uint32_t doneInLogic)
{
    const unsigned grid = streamingGrid(ctx->numTasks);
    const Frame fr = makeFrame(ctx);
    const SceneView sc = makeScene(ctx);
    Timed tm(ctx, FLX_K_MATERIALS);
    if (ctx->params.wfSeparateQueues) // clcontext.cpp:798-812
    {
        // a queue whose BSDF type no uploaded material has stays empty: its kernel is not launched
        const uint32_t have = ctx->materialTypes & ~doneInLogic;
        if (have == 0u)
            return 0;
        if (have & FLX_BXDF_DIFFUSE)
            k_material<FLX_BXDF_DIFFUSE><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_DIFFUSE);
        if (have & FLX_BXDF_GLOSSY)
            k_material<FLX_BXDF_GLOSSY><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_GLOSSY);
        if (have & FLX_BXDF_GGX_ROUGH_REFLECTION)
            k_material<FLX_BXDF_GGX_ROUGH_REFLECTION><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_GGXREFL);
        if (have & FLX_BXDF_GGX_ROUGH_DIELECTRIC)
            k_material<FLX_BXDF_GGX_ROUGH_DIELECTRIC><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_GGXREFR);
        if (have & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC))
            k_material<FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_DELTA);
    }
    else
    {
        constexpr int ALL = FLX_BXDF_DIFFUSE | FLX_BXDF_GLOSSY | FLX_BXDF_GGX_ROUGH_REFLECTION | FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_GGX_ROUGH_DIELECTRIC |
                            FLX_BXDF_IDEAL_DIELECTRIC | FLX_BXDF_EMISSIVE;
        k_material<ALL><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, sc, Q_DIFFUSE);
    }
    return launchCheck(ctx, "k_material");
}

// ---- microkernel integrator (CLContext::enqueueResetKernel ... enqueueSplatPreviewKernel, clcontext.cpp:709-750)
int flx_enqueue_mk_reset(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    Timed tm(ctx, FLX_K_MK_RESET);
    k_mk_reset<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), mk.limit);
    markPixelsWritten(ctx);
    return launchCheck(ctx, "k_mk_reset");
}
FLX_API_CATCH(ctx)

int flx_enqueue_mk_raygen(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    Timed tm(ctx, FLX_K_MK_RAYGEN);
    k_mk_raygen<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params, mk.limit);
    return launchCheck(ctx, "k_mk_raygen");
}
FLX_API_CATCH(ctx)

int flx_enqueue_mk_next_vertex(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    Timed tm(ctx, FLX_K_MK_NEXT_VERTEX);
    if ((rc = launchMkTrace<false>(ctx, mk)))
        return rc;
    k_mk_next_vertex_logic<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), ctx->params, makeScene(ctx), mk);
    return launchCheck(ctx, "k_mk_next_vertex_logic");
}
FLX_API_CATCH(ctx)

int flx_enqueue_mk_sample_bsdf(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    const Frame fr = makeFrame(ctx);
    const SceneView sc = makeScene(ctx);
    Timed tm(ctx, FLX_K_MK_SAMPLE_BSDF);
    const bool nee = ctx->params.sampleExpl && (ctx->params.useEnvMap || ctx->params.useAreaLight);
    CU(cudaMemsetAsync(ctx->mkRayCount, 0, sizeof(uint32_t), ctx->stream));
    CU(cudaMemsetAsync(ctx->mkTypeCounts, 0, MK_NUM_LISTS * sizeof(uint32_t), ctx->stream));
    k_mk_nee_prepare<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk); // light samples (if any) + shading lists
    if ((rc = launchCheck(ctx, "k_mk_nee_prepare")))
        return rc;
    if (nee && (rc = launchMkTrace<true>(ctx, mk)))
        return rc;
    // one shading launch per BSDF type, each over its own list (a warp sees one BSDF)
    const unsigned grid = streamingGrid(mk.limit);
    constexpr int ALL = FLX_BXDF_DIFFUSE | FLX_BXDF_GLOSSY | FLX_BXDF_GGX_ROUGH_REFLECTION | FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_GGX_ROUGH_DIELECTRIC |
                        FLX_BXDF_IDEAL_DIELECTRIC | FLX_BXDF_EMISSIVE;
    const uint32_t have = ctx->materialTypes; // lists of BSDF types the scene does not contain stay empty: not launched
    if (have & FLX_BXDF_DIFFUSE)
        k_mk_shade<FLX_BXDF_DIFFUSE><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_DIFFUSE);
    if (have & FLX_BXDF_GLOSSY)
        k_mk_shade<FLX_BXDF_GLOSSY><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_GLOSSY);
    if (have & FLX_BXDF_GGX_ROUGH_REFLECTION)
        k_mk_shade<FLX_BXDF_GGX_ROUGH_REFLECTION><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_GGX_REFL);
    if (have & FLX_BXDF_GGX_ROUGH_DIELECTRIC)
        k_mk_shade<FLX_BXDF_GGX_ROUGH_DIELECTRIC><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_GGX_REFR);
    if (have & (FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC))
        k_mk_shade<FLX_BXDF_IDEAL_REFLECTION | FLX_BXDF_IDEAL_DIELECTRIC><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_DELTA);
    if (ctx->otherTypes)
        k_mk_shade<ALL><<<grid, FLX_BLOCK, 0, ctx->stream>>>(fr, ctx->params, sc, mk, MK_L_OTHER);
    return launchCheck(ctx, "k_mk_shade");
}
FLX_API_CATCH(ctx)

int flx_enqueue_mk_splat(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    Timed tm(ctx, FLX_K_MK_SPLAT);
    k_mk_splat<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), mk);
    markPixelsWritten(ctx);
    return launchCheck(ctx, "k_mk_splat");
}
FLX_API_CATCH(ctx)

int flx_enqueue_mk_splat_preview(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc || (rc = ensureMk(ctx)))
        return rc;
    CU(cudaSetDevice(ctx->device));
    const MkView mk = makeMk(ctx);
    Timed tm(ctx, FLX_K_MK_SPLAT);
    k_mk_splat_preview<<<mkGrid(mk.limit), FLX_BLOCK, 0, ctx->stream>>>(makeFrame(ctx), mk.limit);
    markPixelsWritten(ctx);
    return launchCheck(ctx, "k_mk_splat_preview");
}
FLX_API_CATCH(ctx)

// Tracer::renderSingle's loop (tracer.cpp:124-150), spp times, no host round trips: camera rays, (maxBounces + 1) x
// (nextVertex, sampleBsdf), splat, display pass.
int flx_render_single(flx_ctx *ctx, uint32_t spp)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc)
        return rc;
    for (uint32_t s = 0; s < spp; s++)
    {
        if ((rc = flx_enqueue_mk_raygen(ctx)))
            return rc;
        for (uint32_t bounce = 0; bounce < ctx->params.maxBounces + 1u; bounce++)
        {
            if ((rc = flx_enqueue_mk_next_vertex(ctx)) || (rc = flx_enqueue_mk_sample_bsdf(ctx)))
                return rc;
        }
        if ((rc = flx_enqueue_mk_splat(ctx)))
            return rc;
        if (ctx->postprocessInLoop && (rc = flx_enqueue_postprocess(ctx)))
            return rc;
    }
    return 0;
}
FLX_API_CATCH(ctx)

static int launchPostprocess(flx_ctx *ctx)
{
    Timed tm(ctx, FLX_K_POSTPROCESS);
    const int all = (ctx->previewStale || !ctx->dirtyPostprocess) ? 1 : 0;
    k_postprocess<<<streamingGrid(ctx->tilePixels), FLX_BLOCK, 0, ctx->cur>>>(reinterpret_cast<const float4 *>(ctx->pixels), reinterpret_cast<float4 *>(ctx->preview),
                                                                            ctx->dirtyPixels, all, ctx->tilePixels, ctx->params.ppParams.exposure,
                                                                            ctx->params.ppParams.tmOperator);
    ctx->previewStale = false;
    if (ctx->denoiser) // mk_postprocess.cl:49-54: the feature buffers go out with the picture
    {
        if (!ctx->aovOut)
            CU(cudaMalloc(&ctx->aovOut, (size_t)ctx->tilePixels * 32));
        float4 *out = reinterpret_cast<float4 *>(ctx->aovOut);
        k_postprocess_aovs<<<streamingGrid(ctx->tilePixels), FLX_BLOCK, 0, ctx->cur>>>(reinterpret_cast<const float4 *>(ctx->denoiserNormal),
                                                                                     reinterpret_cast<const float4 *>(ctx->denoiserAlbedo), out, out + ctx->tilePixels,
                                                                                     ctx->tilePixels);
    }
    return launchCheck(ctx, "k_postprocess");
}

int flx_enqueue_postprocess(flx_ctx *ctx)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    if (!(ctx->overlapPostprocess && ctx->pixelsEventValid))
        return launchPostprocess(ctx);
    // The accumulator has not changed since the event recorded after its last writer (normally this iteration's logic stage),
    // and what was enqueued since (the traversal stages, the counter bookkeeping) does not touch it: start from that event on
    // the third stream, beside that work; the main stream joins before anything enqueued later.
    CU(cudaStreamWaitEvent(ctx->stream3, ctx->evPixels, 0));
    ctx->cur = ctx->stream3;
    rc = launchPostprocess(ctx);
    ctx->cur = ctx->stream;
    if (rc)
        return rc;
    CU(cudaEventRecord(ctx->evPostJoin, ctx->stream3));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->evPostJoin, 0));
    return 0;
}
FLX_API_CATCH(ctx)

// Tracer::useDenoiser (reference: src/kernel_impl.hpp:53, 346, 380, 443 add -DUSE_OPTIX_DENOISER to the logic, nextVertex,
// sampleBsdf and post-process kernels).  The OptiX denoiser itself is out of scope; the feature buffers it consumes are not.
int flx_set_denoiser(flx_ctx *ctx, int enabled)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    ctx->denoiser = enabled != 0;
    return 0;
}
FLX_API_CATCH(ctx)

// which: 0 = first-hit normal, 1 = first-diffuse-hit albedo; processed: 0 = the raw accumulators (sums, w = sample count), 1 = as the
// display pass hands them to the denoiser (divided by w where w > 1; run flx_enqueue_postprocess first)
int flx_read_denoiser_aov(flx_ctx *ctx, int which, int processed, float *rgba, size_t n_pixels)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(rgba != nullptr && (which == 0 || which == 1), "flx_read_denoiser_aov: bad arguments");
    REQUIRE(ctx->pixels && n_pixels <= ctx->tilePixels, "flx_read_denoiser_aov: more pixels requested than the context owns");
    REQUIRE(!processed || ctx->aovOut, "flx_read_denoiser_aov: no display pass has run with the denoiser buffers enabled");
    const float *src = processed ? ctx->aovOut + (which ? (size_t)ctx->tilePixels * 4 : 0) : (which ? ctx->denoiserAlbedo : ctx->denoiserNormal);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(rgba, src, n_pixels * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_read_preview(flx_ctx *ctx, float *rgba, size_t n_pixels)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(rgba != nullptr, "flx_read_preview: null destination");
    REQUIRE(ctx->preview && n_pixels <= ctx->tilePixels, "flx_read_preview: more pixels requested than the context owns");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(rgba, ctx->preview, n_pixels * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_enqueue_clear_queues(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemsetAsync(ctx->counters, 0, sizeof(flx_QueueCounters), ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_enqueue_get_counters(flx_ctx *ctx, flx_QueueCounters *host_out)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(host_out != nullptr, "flx_enqueue_get_counters: null destination");
    CU(cudaSetDevice(ctx->device));
    if ((int)ctx->pendingCounterReads.size() >= flx_ctx::kCounterRing)
    {
        int rc = flx_finish(ctx);
        if (rc)
            return rc;
    }
    const int slot = ctx->counterRingPos;
    ctx->counterRingPos = (ctx->counterRingPos + 1) % flx_ctx::kCounterRing;
    CU(cudaMemcpyAsync(ctx->pinnedCounters + slot, ctx->counters, sizeof(flx_QueueCounters), cudaMemcpyDeviceToHost, ctx->stream));
    ctx->pendingCounterReads.push_back({slot, host_out});
    return 0;
}
FLX_API_CATCH(ctx)

int flx_finish(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gatherInFlight) // the caller sees ONE in-order queue: a gather enqueued before this call is complete after it
    {
        CU(cudaEventSynchronize(ctx->evGatherDone));
        ctx->gatherInFlight = false;
    }
    for (auto &r : ctx->pendingCounterReads)
        *r.second = ctx->pinnedCounters[r.first];
    ctx->pendingCounterReads.clear();
    drainEvents(ctx);
    return 0;
}
FLX_API_CATCH(ctx)

int flx_update_pixel_index(flx_ctx *ctx, uint32_t num_pixels, uint32_t num_new_paths)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(num_pixels > 0, "flx_update_pixel_index: zero pixels");
    CU(cudaSetDevice(ctx->device));
    // reference: host-tracked index, advanced and written with a NON-blocking 4-byte copy (clcontext.cpp:891-895).  Same here;
    // only after flx_render (which advances the device copy itself) the host value is refreshed first.
    if (ctx->pixelIdxAdvancedOnDevice)
    {
        CU(cudaMemcpyAsync(&ctx->hostPixelIdx, ctx->currPixelIdx, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->pixelIdxAdvancedOnDevice = false;
    }
    ctx->hostPixelIdx = (ctx->hostPixelIdx + num_new_paths) % num_pixels;
    if (ctx->pixelIdxRingPos == flx_ctx::kCounterRing) // every staging slot may still be in flight: drain before reusing them
    {
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->pixelIdxRingPos = 0;
    }
    uint32_t *slot = ctx->pinnedPixelIdx + ctx->pixelIdxRingPos++;
    *slot = ctx->hostPixelIdx;
    CU(cudaMemcpyAsync(ctx->currPixelIdx, slot, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_reset_pixel_index(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    ctx->hostPixelIdx = 0;
    ctx->pixelIdxAdvancedOnDevice = false;
    CU(cudaMemsetAsync(ctx->currPixelIdx, 0, sizeof(uint32_t), ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_render(flx_ctx *ctx, uint32_t n_iterations)
try
{
    int rc = checkReady(ctx, true, true);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    const IterationState it = makeIter(ctx);
    if (n_iterations)
        ctx->pixelIdxAdvancedOnDevice = true;
    for (uint32_t i = 0; i < n_iterations; i++) // tracer.cpp:433-439, 465
    {
        if (ctx->fuseStages) // logic + raygen + materials in one pass over the path state; same state, queues and counters
        {
            if ((rc = launchLogic(ctx, 0, true)))
                return rc;
        }
        else
        {
            if ((rc = flx_enqueue_logic(ctx, 0)))
                return rc;
            if ((rc = flx_enqueue_raygen(ctx)))
                return rc;
            if ((rc = flx_enqueue_materials(ctx)))
                return rc;
        }
        // The display pass reads the accumulator, which only the logic stage writes (splat): the picture after this iteration is
        // fixed from here on, so the pass can run beside the two traversal stages instead of after them (tracer.cpp:447 runs it
        // last; same input, same output).  Joined before the next iteration's logic may splat again.  Level 2 only: in this
        // loop it gains 0.4 % and makes the extension kernel's own elapsed time (the roofline's denominator) meaningless.
        const bool postBeside = ctx->postprocessInLoop && ctx->overlapPostprocess == 2;
        if (postBeside)
        {
            CU(cudaEventRecord(ctx->evPostFork, ctx->stream));
            CU(cudaStreamWaitEvent(ctx->stream3, ctx->evPostFork, 0));
            ctx->cur = ctx->stream3;
            rc = launchPostprocess(ctx);
            ctx->cur = ctx->stream;
            if (rc)
                return rc;
            CU(cudaEventRecord(ctx->evPostJoin, ctx->stream3));
        }
        // (no counter snapshot here: nothing between this point and k_end_iteration changes the queue counters -- the traversal stages only
        //  read them -- so the end-of-iteration kernel reads them in place)
        // extension then shadow rays: the second call overlaps the first on a second stream (see flx_enqueue_shadowrays)
        if ((rc = flx_enqueue_extrays(ctx)))
            return rc;
        if ((rc = flx_enqueue_shadowrays(ctx)))
            return rc;
        {
            Timed tm(ctx, FLX_K_END_ITERATION);
            IterationState itEnd = it;
            itEnd.snapshot = it.counters;
            itEnd.fetch = ctx->fetchCounters;
            itEnd.scanTiles = ctx->scanTiles;
            itEnd.nScanTiles = ctx->lastLogicTiles;
            itEnd.scanTicket = ctx->scanTicket;
            k_end_iteration<<<1, 256, 0, ctx->stream>>>(itEnd);
            ctx->fetchClean[0] = ctx->fetchClean[1] = true;
            ctx->scanClean = true;
        }
        if ((rc = launchCheck(ctx, "k_end_iteration")))
            return rc;
        if (postBeside)
            CU(cudaStreamWaitEvent(ctx->stream, ctx->evPostJoin, 0));
        else if (ctx->postprocessInLoop && (rc = flx_enqueue_postprocess(ctx))) // tracer.cpp:447
            return rc;
    }
    return 0;
}
FLX_API_CATCH(ctx)

int flx_render_timed(flx_ctx *ctx, uint32_t n_iterations, float *elapsed_ms)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(elapsed_ms != nullptr, "flx_render_timed: null destination");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(ctx->evStart, ctx->stream));
    int rc = flx_render(ctx, n_iterations);
    if (rc)
        return rc;
    CU(cudaEventRecord(ctx->evStop, ctx->stream));
    CU(cudaEventSynchronize(ctx->evStop));
    CU(cudaEventElapsedTime(elapsed_ms, ctx->evStart, ctx->evStop));
    drainEvents(ctx);
    return 0;
}
FLX_API_CATCH(ctx)

int flx_timer_begin(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaEventRecord(ctx->evStart, ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_timer_end(flx_ctx *ctx, float *elapsed_ms)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(elapsed_ms != nullptr, "flx_timer_end: null destination");
    CU(cudaSetDevice(ctx->device));
    if (ctx->gatherInFlight) // a gather still running on its own stream belongs to the timed work
        CU(cudaStreamWaitEvent(ctx->stream, ctx->evGatherDone, 0));
    CU(cudaEventRecord(ctx->evStop, ctx->stream));
    CU(cudaEventSynchronize(ctx->evStop));
    CU(cudaEventElapsedTime(elapsed_ms, ctx->evStart, ctx->evStop));
    drainEvents(ctx);
    return 0;
}
FLX_API_CATCH(ctx)

int flx_set_tuning(flx_ctx *ctx, int key, int value)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    switch (key)
    {
    case FLX_TUNE_TRACE_VARIANT:
        REQUIRE(value >= 0 && value <= 3, "flx_set_tuning: trace variant must be 0, 1, 2 or 3");
        ctx->traceVariant = value;
        return 0;
    case FLX_TUNE_BVH_TRI_COST:
        REQUIRE(value >= 25 && value <= 1600, "flx_set_tuning: triangle cost must be in 25..1600 percent");
        ctx->bvhTriCostPercent = value;
        return 0;
    case FLX_TUNE_BVH_DEPTH_LIMIT:
        REQUIRE(value >= 1 && value <= 62, "flx_set_tuning: hierarchy depth limit must be in 1..62");
        ctx->bvhDepthLimit = value;
        return 0;
    case FLX_TUNE_MATERIAL_MASK:
        ctx->useMaterialMask = value != 0;
        return 0;
    case FLX_TUNE_BVH_REINSERT:
        REQUIRE(value >= 0 && value <= 64, "flx_set_tuning: reinsertion iterations must be in 0..64");
        ctx->bvhReinsertIterations = value;
        return 0;
    case FLX_TUNE_SHADOW_LEFT_FIRST:
    case FLX_TUNE_PREFETCH_CHILDREN:
        // measured slower in rounds 1 / 2 (profiles/r2_knob_sweep.txt) and removed from the traversal loop, where even a skipped branch costs
        REQUIRE(value == 0, "flx_set_tuning: this experiment was measured, rejected and removed (DESIGN.md 4.1); only 0 is accepted");
        return 0;
    case FLX_TUNE_LOGIC_TILE:
        REQUIRE(value == 128 || value == 256, "flx_set_tuning: logic tile must be 128 or 256");
        ctx->logicTile = value;
        return 0;
    case FLX_TUNE_GATHER_DIRECT:
        ctx->gatherDirect = value != 0;
        return 0;
    case FLX_TUNE_GATHER_PRIORITY:
        REQUIRE(ctx->gatherStream == nullptr, "flx_set_tuning: the gather stream already exists (set its priority before the first flx_gather_pixels)");
        ctx->gatherPriority = value != 0;
        return 0;
    case FLX_TUNE_INNER_BIAS:
        REQUIRE(value >= -32 && value <= 32, "flx_set_tuning: inner bias must be in -32..32");
        ctx->innerBias = value;
        return 0;
    case FLX_TUNE_FETCH_THRESHOLD:
        REQUIRE(value >= 1 && value <= 32, "flx_set_tuning: fetch threshold must be in 1..32");
        ctx->fetchThreshold = value;
        return 0;
    case FLX_TUNE_INNER_MIN:
        REQUIRE(value >= 1 && value <= 32, "flx_set_tuning: inner-phase minimum must be in 1..32");
        ctx->innerMin = value;
        return 0;
    case FLX_TUNE_EXT_MIN_BLOCKS:
    case FLX_TUNE_SHADOW_MIN_BLOCKS:
        REQUIRE(value == 8 || value == 9 || value == 10, "flx_set_tuning: trace min blocks must be 8, 9 or 10");
        (key == FLX_TUNE_EXT_MIN_BLOCKS ? ctx->extMinBlocks : ctx->shadowMinBlocks) = value;
        return 0;
    case FLX_TUNE_MAX_L1:
        ctx->maxL1 = value != 0;
        return 0;
    case FLX_TUNE_SMEM_STACK:
        REQUIRE(value == -1 || value == 0 || value == 1 || value == 4 || value == 8 || value == 24, "flx_set_tuning: smem stack levels must be 0, 4, 8 or 24 (1 = 24), or -1 for local memory + the newest entry in a register");
        ctx->smemStack = value == 1 ? 24 : value;
        return 0;
    case FLX_TUNE_POSTPROCESS_IN_LOOP:
        ctx->postprocessInLoop = value != 0;
        return 0;
    case FLX_TUNE_OVERLAP_TRACE:
        ctx->overlapTrace = value != 0;
        return 0;
    case FLX_TUNE_DIRTY_POSTPROCESS:
        ctx->dirtyPostprocess = value != 0;
        return 0;
    case FLX_TUNE_OVERLAP_POSTPROCESS:
        REQUIRE(value >= 0 && value <= 2, "flx_set_tuning: display-pass overlap must be 0, 1 or 2");
        ctx->overlapPostprocess = value;
        return 0;
    case FLX_TUNE_FETCH_CHUNK:
        REQUIRE(value >= 32 && value <= 4096, "flx_set_tuning: fetch chunk must be in 32..4096");
        ctx->fetchChunk = value;
        return 0;
    case FLX_TUNE_LOGIC_MIN_BLOCKS:
        REQUIRE(value == 2 || value == 3 || value == 4, "flx_set_tuning: logic min blocks must be 2, 3 or 4");
        ctx->logicMinBlocks = value;
        return 0;
    case FLX_TUNE_L2_PERSIST:
        REQUIRE(value >= 0 && value <= 2, "flx_set_tuning: L2 persistence must be 0, 1 (TTri) or 2 (TNode)");
        ctx->l2Persist = value;
        CU(cudaSetDevice(ctx->device));
        return applyL2Persist(ctx);
    case FLX_TUNE_REPACK_ON_HOST:
        ctx->repackOnHost = value != 0;
        return 0;
    case FLX_TUNE_FUSE_STAGES:
        ctx->fuseStages = value != 0;
        return 0;
    case FLX_TUNE_FUSED_MIN_BLOCKS:
        REQUIRE(value >= 1 && value <= 4, "flx_set_tuning: fused min blocks must be 1..4");
        ctx->fusedMinBlocks = value;
        return 0;
    case FLX_TUNE_TOP_NODES:
        REQUIRE(value >= 0 && value <= 4096, "flx_set_tuning: top nodes must be in 0..4096");
        ctx->topNodes = value;
        return 0;
    case FLX_TUNE_TRACE_BLOCKS_PER_SM:
        REQUIRE(value >= 0 && value <= 32, "flx_set_tuning: blocks per SM must be in 0..32");
        ctx->traceBlocksPerSM = value;
        return 0;
    }
    return fail(ctx, FLX_E_INVALID, "flx_set_tuning: unknown key %d", key);
}
FLX_API_CATCH(ctx)

int flx_set_counting(flx_ctx *ctx, int enabled)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    ctx->counting = enabled != 0;
    if (ctx->counting)
        CU(cudaMemsetAsync(ctx->traceCounts, 0, 10 * sizeof(unsigned long long), ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_get_trace_counts(flx_ctx *ctx, flx_TraceCounts *ext, flx_TraceCounts *shadow)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(ext && shadow, "flx_get_trace_counts: null destination");
    CU(cudaSetDevice(ctx->device));
    unsigned long long h[10];
    CU(cudaMemcpyAsync(h, ctx->traceCounts, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    flx_TraceCounts *o[2] = {ext, shadow};
    for (int k = 0; k < 2; k++)
    {
        o[k]->nodes = h[5 * k + 0];
        o[k]->boxes = h[5 * k + 1];
        o[k]->tris = h[5 * k + 2];
        o[k]->updates = h[5 * k + 3];
        o[k]->rays = h[5 * k + 4];
    }
    return 0;
}
FLX_API_CATCH(ctx)

int flx_reset_stats(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gatherStream)
        CU(cudaStreamSynchronize(ctx->gatherStream));
    drainEvents(ctx);
    CU(cudaMemsetAsync(ctx->stats, 0, sizeof(flx_RenderStats64), ctx->stream));
    for (int k = 0; k < FLX_K_COUNT; k++)
    {
        ctx->kernelMs[k] = 0.0;
        ctx->kernelLaunches[k] = 0;
    }
    return 0;
}
FLX_API_CATCH(ctx)

int flx_get_stats(flx_ctx *ctx, flx_RenderStats64 *out)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(out != nullptr, "flx_get_stats: null destination");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(out, ctx->stats, sizeof *out, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_set_profiling(flx_ctx *ctx, int enabled)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    ctx->profiling = enabled != 0;
    return 0;
}
FLX_API_CATCH(ctx)

int flx_get_kernel_ms(flx_ctx *ctx, int kernel_id, float *total_ms, uint32_t *launches)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(kernel_id >= 0 && kernel_id < FLX_K_COUNT, "flx_get_kernel_ms: bad kernel id");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    if (ctx->gatherStream)
        CU(cudaStreamSynchronize(ctx->gatherStream));
    drainEvents(ctx);
    if (total_ms)
        *total_ms = (float)ctx->kernelMs[kernel_id];
    if (launches)
        *launches = ctx->kernelLaunches[kernel_id];
    return 0;
}
FLX_API_CATCH(ctx)

int flx_read_pixels(flx_ctx *ctx, float *rgba, size_t n_pixels)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(rgba != nullptr, "flx_read_pixels: null destination");
    REQUIRE(ctx->pixels && n_pixels <= ctx->tilePixels, "flx_read_pixels: more pixels requested than the context owns");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(rgba, ctx->pixels, n_pixels * 4 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

// CLContext::saveImage (clcontext.hpp:78; clcontext.cpp:386-465): *.hdr from the accumulator, anything else from the preview
int flx_save_image(flx_ctx *ctx, const char *filename)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    REQUIRE(filename != nullptr, "flx_save_image: null file name");
    const bool tiled = ctx->nParts != 1;
    REQUIRE(!tiled || (ctx->comm != nullptr && ctx->lastGatherRoot == ctx->rank),
            "flx_save_image: this context renders a tile; gather the full image first (flx_gather_pixels) and save it on the root");
    const std::string name(filename);
    const bool hdr = name.size() >= 4 && (name.compare(name.size() - 4, 4, ".hdr") == 0 || name.compare(name.size() - 4, 4, ".HDR") == 0);
    const size_t count = tiled ? (size_t)ctx->width * ctx->height : (size_t)ctx->tilePixels;
    std::vector<float> host(count * 4);
    if (tiled) // the root of the last gather writes the gathered frame: accumulators for .hdr, their display pass otherwise
        rc = flx_read_gathered(ctx, hdr ? 0 : 1, host.data(), count);
    else
        rc = hdr ? flx_read_pixels(ctx, host.data(), ctx->tilePixels) : flx_read_preview(ctx, host.data(), ctx->tilePixels);
    if (rc)
        return rc;
    if (flx_write_image(filename, host.data(), ctx->width, ctx->height) != 0)
        return fail(ctx, FLX_E_INVALID, "flx_save_image: %s", flx_io_last_error());
    return 0;
}
FLX_API_CATCH(ctx)

// test/diagnostic: the traversal layout as it lives on the device (TNode and TTri records of 16 floats each)
int flx_read_traversal_layout(flx_ctx *ctx, float *tnodes_out, uint32_t *n_tnodes, float *ttris_out, uint32_t *n_ttris, int32_t *root_ref)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(n_tnodes && n_ttris, "flx_read_traversal_layout: null count");
    REQUIRE(ctx->sceneReady, "flx_read_traversal_layout: no scene uploaded");
    CU(cudaSetDevice(ctx->device));
    if (tnodes_out && ttris_out)
    {
        REQUIRE(*n_tnodes >= ctx->nTNodes && *n_ttris >= ctx->nTTris, "flx_read_traversal_layout: arrays too small");
        CU(cudaMemcpyAsync(tnodes_out, ctx->tnodes, (size_t)ctx->nTNodes * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ttris_out, ctx->ttris, (size_t)ctx->nTTris * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    *n_tnodes = ctx->nTNodes;
    *n_ttris = ctx->nTTris;
    if (root_ref)
        *root_ref = ctx->rootRef;
    return 0;
}
FLX_API_CATCH(ctx)

int flx_read_tasks(flx_ctx *ctx, uint32_t *slots_out)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(slots_out != nullptr, "flx_read_tasks: null destination");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(slots_out, ctx->tasks, (size_t)ctx->numTasks * FLX_NUM_SLOTS * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_write_tasks(flx_ctx *ctx, const uint32_t *slots_in)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(slots_in != nullptr, "flx_write_tasks: null source");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->tasks, slots_in, (size_t)ctx->numTasks * FLX_NUM_SLOTS * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_read_queue(flx_ctx *ctx, int queue_id, uint32_t *out, uint32_t max_entries)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(queue_id >= 0 && queue_id < 8 && out, "flx_read_queue: bad arguments");
    CU(cudaSetDevice(ctx->device));
    const uint32_t n = std::min(max_entries, ctx->numTasks);
    CU(cudaMemcpyAsync(out, ctx->queues[queue_id], (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_write_queue(flx_ctx *ctx, int queue_id, const uint32_t *entries, uint32_t n)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(queue_id >= 0 && queue_id < 8 && (entries || n == 0) && n <= ctx->numTasks, "flx_write_queue: bad arguments");
    CU(cudaSetDevice(ctx->device));
    if (n)
        CU(cudaMemcpyAsync(ctx->queues[queue_id], entries, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

int flx_write_counters(flx_ctx *ctx, const flx_QueueCounters *in)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(in != nullptr, "flx_write_counters: null source");
    CU(cudaSetDevice(ctx->device));
    CU(cudaMemcpyAsync(ctx->counters, in, sizeof *in, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return 0;
}
FLX_API_CATCH(ctx)

// ---- checkpoint / resume (new; SURVEY 5: the reference can only restart a render from scratch -- its caches hold the
// hierarchy, the camera state and compiled kernels, never the accumulator).  One file holds everything an interrupted
// wavefront or microkernel render needs to continue bit-identically: path state, queues, counters, pixel index, statistics
// and the accumulator.  Scene, environment map and params are NOT in it: upload them as usual, then load.
namespace
{
struct CheckpointHeader
{
    char magic[8]; // "FLXCKPT2"
    uint32_t numTasks, width, height, tilePixels, part, nParts, stripeRows, hostPixelIdx;
    uint64_t sceneHash, paramsHash; // what the path state was computed FOR: resuming against anything else gives a wrong picture
};

uint64_t fnv1a(uint64_t h, const void *data, size_t bytes)
{
    const unsigned char *p = static_cast<const unsigned char *>(data);
    for (size_t i = 0; i < bytes; i++)
        h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

// everything in RenderParams that decides the accumulator (the display-only ppParams do not)
uint64_t paramsFingerprint(const flx_RenderParams &p)
{
    flx_RenderParams q = p;
    memset(&q.ppParams, 0, sizeof q.ppParams);
    return fnv1a(14695981039346656037ull, &q, sizeof q);
}
} // namespace

int flx_checkpoint_save(flx_ctx *ctx, const char *path)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    REQUIRE(path != nullptr, "flx_checkpoint_save: null path");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    CheckpointHeader h;
    memset(&h, 0, sizeof h);
    memcpy(h.magic, "FLXCKPT2", 8);
    h.numTasks = ctx->numTasks; h.width = ctx->width; h.height = ctx->height; h.tilePixels = ctx->tilePixels;
    h.part = ctx->part; h.nParts = ctx->nParts; h.stripeRows = ctx->stripeRows;
    h.sceneHash = ctx->sceneHash;
    h.paramsHash = paramsFingerprint(ctx->params);
    CU(cudaMemcpyAsync(&h.hostPixelIdx, ctx->currPixelIdx, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    FILE *fp = std::fopen(path, "wb");
    if (!fp)
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_save: cannot create %s", path);
    bool ok = std::fwrite(&h, sizeof h, 1, fp) == 1;
    std::vector<unsigned char> buf;
    auto dump = [&](const void *dev, size_t bytes) {
        buf.resize(bytes);
        if (cudaMemcpyAsync(buf.data(), dev, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess || cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            ok = false;
        ok = ok && std::fwrite(buf.data(), 1, bytes, fp) == bytes;
    };
    dump(ctx->tasks, (size_t)ctx->numTasks * FLX_NUM_SLOTS * 4);
    for (int q = 0; q < 8; q++)
        dump(ctx->queues[q], (size_t)ctx->numTasks * 4);
    dump(ctx->counters, sizeof(flx_QueueCounters));
    dump(ctx->stats, sizeof(flx_RenderStats64));
    dump(ctx->pixels, (size_t)ctx->tilePixels * 16);
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok)
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_save: write error on %s", path);
    return 0;
}
FLX_API_CATCH(ctx)

int flx_checkpoint_load(flx_ctx *ctx, const char *path)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    REQUIRE(path != nullptr, "flx_checkpoint_load: null path");
    CU(cudaSetDevice(ctx->device));
    CU(cudaStreamSynchronize(ctx->stream));
    FILE *fp = std::fopen(path, "rb");
    if (!fp)
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: cannot open %s", path);
    CheckpointHeader h;
    if (std::fread(&h, sizeof h, 1, fp) != 1 || memcmp(h.magic, "FLXCKPT2", 8) != 0)
    {
        std::fclose(fp);
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: %s is not a checkpoint (or one of an older format)", path);
    }
    if (h.numTasks != ctx->numTasks || h.width != ctx->width || h.height != ctx->height || h.tilePixels != ctx->tilePixels || h.part != ctx->part ||
        h.nParts != ctx->nParts || h.stripeRows != ctx->stripeRows)
    {
        std::fclose(fp);
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: checkpoint is for %u paths, %ux%u, tile %u/%u; this context has %u paths, %ux%u, tile %u/%u", h.numTasks,
                    h.width, h.height, h.part, h.nParts, ctx->numTasks, ctx->width, ctx->height, ctx->part, ctx->nParts);
    }
    if (h.sceneHash != ctx->sceneHash || h.paramsHash != paramsFingerprint(ctx->params))
    {
        std::fclose(fp);
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: the checkpoint was written for a different %s; its path state would continue into a wrong picture",
                    h.sceneHash != ctx->sceneHash ? "scene" : "camera / light / sampling setup (RenderParams)");
    }
    // Read and VALIDATE everything on the host before the first byte reaches the device: the kernels index the path state with
    // queue entries and the accumulator with the pixelIndex slot, so a damaged file must not get that far.
    const size_t n = ctx->numTasks;
    std::vector<uint32_t> tasks(n * FLX_NUM_SLOTS), queues(n * 8);
    flx_QueueCounters counters;
    flx_RenderStats64 stats;
    std::vector<float> pixels((size_t)ctx->tilePixels * 4);
    bool ok = std::fread(tasks.data(), 4, tasks.size(), fp) == tasks.size() && std::fread(queues.data(), 4, queues.size(), fp) == queues.size() &&
              std::fread(&counters, sizeof counters, 1, fp) == 1 && std::fread(&stats, sizeof stats, 1, fp) == 1 &&
              std::fread(pixels.data(), 4, pixels.size(), fp) == pixels.size();
    std::fclose(fp);
    if (!ok)
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: %s is truncated", path);
    if (h.hostPixelIdx >= ctx->tilePixels)
        return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: pixel index %u outside the %u-pixel image", h.hostPixelIdx, ctx->tilePixels);
    const uint32_t *cnt = reinterpret_cast<const uint32_t *>(&counters);
    for (int q = 0; q < 8; q++)
    {
        if (cnt[q] > n)
            return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: queue %d holds %u entries, more than the %zu paths", q, cnt[q], n);
        for (uint32_t k = 0; k < cnt[q]; k++)
            if (queues[(size_t)q * n + k] >= n)
                return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: queue %d entry %u names path %u of %zu", q, k, queues[(size_t)q * n + k], n);
    }
    for (size_t g = 0; g < n; g++)
        if (tasks[(size_t)FLX_S_PIXEL_INDEX * n + g] >= ctx->tilePixels)
            return fail(ctx, FLX_E_INVALID, "flx_checkpoint_load: path %zu splats into pixel %u of %u", g, tasks[(size_t)FLX_S_PIXEL_INDEX * n + g], ctx->tilePixels);
    // on the context's stream (the work streams do not order against the legacy default stream), one synchronisation at the end
    CU(cudaMemcpyAsync(ctx->tasks, tasks.data(), tasks.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    for (int q = 0; q < 8; q++)
        CU(cudaMemcpyAsync(ctx->queues[q], queues.data() + (size_t)q * n, n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->counters, &counters, sizeof counters, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->stats, &stats, sizeof stats, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->pixels, pixels.data(), pixels.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->currPixelIdx, &h.hostPixelIdx, sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->hostPixelIdx = h.hostPixelIdx;
    ctx->pixelIdxAdvancedOnDevice = false;
    ctx->previewStale = true;
    markPixelsWritten(ctx);
    return 0;
}
FLX_API_CATCH(ctx)

// ---- NCCL gather of the per-tile radiance buffers: the only collective of the path
int flx_comm_unique_id(void *out128)
try
{
    flx_ctx tmp;
    flx_ctx *ctx = &tmp;
    int rc = loadNccl(ctx);
    if (rc)
    {
        g_create_error = tmp.error;
        return rc;
    }
    int r = tmp.nccl.GetUniqueId(out128);
    if (r != 0)
    {
        g_create_error = std::string("ncclGetUniqueId: ") + tmp.nccl.GetErrorString(r);
        return FLX_E_NCCL;
    }
    return 0;
}
FLX_API_CATCH((flx_ctx *)nullptr)

int flx_comm_init(flx_ctx *ctx, const void *unique_id128, int rank, int nranks)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    TOUCH(ctx);
    REQUIRE(unique_id128 && nranks > 0 && rank >= 0 && rank < nranks, "flx_comm_init: bad arguments");
    int rc = loadNccl(ctx);
    if (rc)
        return rc;
    CU(cudaSetDevice(ctx->device));
    Id128 id;
    memcpy(&id, unique_id128, sizeof id);
    int r = ctx->nccl.CommInitRank(&ctx->comm, nranks, id, rank);
    if (r != 0)
        return fail(ctx, FLX_E_NCCL, "ncclCommInitRank: %s", ctx->nccl.GetErrorString(r));
    ctx->rank = rank;
    ctx->nranks = nranks;
    return 0;
}
FLX_API_CATCH(ctx)

int flx_comm_destroy(flx_ctx *ctx)
try
{
    if (!ctx)
        return FLX_E_INVALID;
    if (ctx->comm && ctx->nccl.CommDestroy)
        ctx->nccl.CommDestroy(ctx->comm);
    ctx->comm = nullptr;
    return 0;
}
FLX_API_CATCH(ctx)

int flx_gather_pixels(flx_ctx *ctx, int root, float *full_rgba_host_or_null)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    REQUIRE(ctx->comm != nullptr, "flx_gather_pixels: flx_comm_init has not been called");
    REQUIRE(ctx->nranks == (int)ctx->nParts && ctx->rank == (int)ctx->part, "flx_gather_pixels: tile partition does not match the communicator");
    REQUIRE(root >= 0 && root < ctx->nranks, "flx_gather_pixels: bad root");
    CU(cudaSetDevice(ctx->device));
    // every rank's tile has at most maxTile pixels; ragged tiles are sent at their own size
    uint32_t maxTile = 0;
    std::vector<uint32_t> tileOf(ctx->nranks);
    for (int r = 0; r < ctx->nranks; r++)
    {
        tileOf[r] = localRows(ctx->height, r, ctx->nParts, ctx->stripeRows) * ctx->width;
        maxTile = std::max(maxTile, tileOf[r]);
    }
    const size_t fullPixels = (size_t)ctx->width * ctx->height;
    // Two ways to land the tiles in the full image on the root:
    //   staged (default): one send / recv per rank into a rank-major buffer, then k_deinterleave.
    //   direct (FLX_TUNE_GATHER_DIRECT = 1): stripe by stripe, sender and root issue one ncclSend / ncclRecv per stripe and the root
    //     receives each stripe straight into its rows of the full image (a stripe of stripeRows rows is contiguous in both layouts); the
    //     root's own stripes are one strided device copy.  No gather buffer, no de-interleave pass -- and 3-4x SLOWER: a grouped NCCL
    //     point-to-point operation costs ~14 us whatever its size, and there is one per stripe (135 at 3840x2160 on 2 GPUs).  Kept,
    //     bit-identical (tests run both), as the measured reason why the gather stays one transfer per rank.
    const bool direct = ctx->gatherDirect != 0;
    const bool isRoot = ctx->rank == root;
    const bool needSnapshot = !(direct && isRoot); // the root's direct path copies its stripes into the full image instead
    const bool grow = (isRoot && ((!direct && ctx->gatherBufPixels < (size_t)maxTile * ctx->nranks) || ctx->fullImagePixels < fullPixels)) ||
                      (needSnapshot && ctx->gatherSnapshotPixels < ctx->tilePixels);
    if (grow && ctx->gatherStream) // a gather still in flight uses the buffers about to be replaced
        CU(cudaStreamSynchronize(ctx->gatherStream));
    if (isRoot && !direct && ctx->gatherBufPixels < (size_t)maxTile * ctx->nranks)
    {
        freeDev(ctx->gatherBuf);
        ctx->gatherBufPixels = 0;
        CU(cudaMalloc(&ctx->gatherBuf, (size_t)maxTile * ctx->nranks * 16));
        ctx->gatherBufPixels = (size_t)maxTile * ctx->nranks;
    }
    if (isRoot && ctx->fullImagePixels < fullPixels)
    {
        freeDev(ctx->fullImage);
        ctx->fullImagePixels = 0;
        CU(cudaMalloc(&ctx->fullImage, fullPixels * 16));
        ctx->fullImagePixels = fullPixels;
    }
    if (needSnapshot && ctx->gatherSnapshotPixels < ctx->tilePixels)
    {
        freeDev(ctx->gatherSnapshot);
        ctx->gatherSnapshotPixels = 0;
        CU(cudaMalloc(&ctx->gatherSnapshot, (size_t)ctx->tilePixels * 16 * 2));
        ctx->gatherSnapshotPixels = ctx->tilePixels;
        ctx->snapshotBusy[0] = ctx->snapshotBusy[1] = false;
    }
    if (!ctx->gatherStream)
    {
        int prioLow = 0, prioHigh = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prioLow, &prioHigh));
        // Low priority by default: the render kernels are persistent and fill every SM, so NCCL's few CTAs get their slots when a
        // traversal kernel's CTAs retire; with the high priority they are placed ahead of the next render kernel's CTAs instead.
        CU(cudaStreamCreateWithPriority(&ctx->gatherStream, cudaStreamNonBlocking, ctx->gatherPriority ? prioHigh : prioLow));
        CU(cudaEventCreateWithFlags(&ctx->evSnapshot, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->evGatherDone, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->evSnapshotFree[0], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&ctx->evSnapshotFree[1], cudaEventDisableTiming));
    }
    // the frame as of now: a snapshot on the render stream (ordered after the last splat, before the next) into the buffer
    // the gather before the previous one used -- so the render stream only ever waits for a gather two frames old
    const int par = ctx->gatherParity;
    const size_t stripeFloats = (size_t)ctx->stripeRows * ctx->width * 4; // one full stripe
    float *snapshot = nullptr;
    if (needSnapshot)
    {
        ctx->gatherParity ^= 1;
        snapshot = ctx->gatherSnapshot + (size_t)par * ctx->gatherSnapshotPixels * 4;
        if (ctx->snapshotBusy[par])
            CU(cudaStreamWaitEvent(ctx->stream, ctx->evSnapshotFree[par], 0));
        CU(cudaMemcpyAsync(snapshot, ctx->pixels, (size_t)ctx->tilePixels * 16, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    else
    {
        // root, direct: its stripes go from the accumulator into their rows of the full image in one strided copy (local stripe k is
        // global stripe k * nParts + part; a ragged last stripe is simply fewer bytes).  The full image's other rows belong to the
        // receives of this and earlier gathers, which are ordered on the gather stream.
        const uint32_t myRows = localRows(ctx->height, ctx->part, ctx->nParts, ctx->stripeRows);
        const uint32_t fullStripes = myRows / ctx->stripeRows, tailRows = myRows % ctx->stripeRows;
        float *dst0 = ctx->fullImage + (size_t)ctx->part * stripeFloats;
        if (fullStripes)
            CU(cudaMemcpy2DAsync(dst0, stripeFloats * ctx->nParts * sizeof(float), ctx->pixels, stripeFloats * sizeof(float), stripeFloats * sizeof(float), fullStripes,
                                 cudaMemcpyDeviceToDevice, ctx->stream));
        if (tailRows)
            CU(cudaMemcpyAsync(dst0 + (size_t)fullStripes * ctx->nParts * stripeFloats, ctx->pixels + (size_t)fullStripes * stripeFloats, (size_t)tailRows * ctx->width * 16,
                               cudaMemcpyDeviceToDevice, ctx->stream));
    }
    CU(cudaEventRecord(ctx->evSnapshot, ctx->stream));
    CU(cudaStreamWaitEvent(ctx->gatherStream, ctx->evSnapshot, 0));
    ctx->cur = ctx->gatherStream;
    {
    Timed tm(ctx, FLX_K_GATHER); // NCCL send/recv group (+ de-interleave when staged), on the gather stream
    const int ncclFloat = 7; // ncclFloat32
    int r = ctx->nccl.GroupStart();
    if (direct)
    {
        // stripe s (global) belongs to rank s % nParts and is local stripe s / nParts there; same order on both sides of each pair
        const uint32_t nStripes = (ctx->height + ctx->stripeRows - 1) / ctx->stripeRows;
        for (uint32_t s = 0; s < nStripes && r == 0; s++)
        {
            const int owner = (int)(s % ctx->nParts);
            const uint32_t rows = std::min(ctx->stripeRows, ctx->height - s * ctx->stripeRows);
            const size_t count = (size_t)rows * ctx->width * 4;
            if (owner == root)
                continue;
            if (ctx->rank == owner)
                r = ctx->nccl.Send(snapshot + (size_t)(s / ctx->nParts) * stripeFloats, count, ncclFloat, root, ctx->comm, ctx->gatherStream);
            else if (isRoot)
                r = ctx->nccl.Recv(ctx->fullImage + (size_t)s * stripeFloats, count, ncclFloat, owner, ctx->comm, ctx->gatherStream);
        }
    }
    else
    {
        if (r == 0)
            r = ctx->nccl.Send(snapshot, (size_t)ctx->tilePixels * 4, ncclFloat, root, ctx->comm, ctx->gatherStream);
        if (r == 0 && isRoot)
            for (int src = 0; src < ctx->nranks && r == 0; src++)
                r = ctx->nccl.Recv(ctx->gatherBuf + (size_t)src * maxTile * 4, (size_t)tileOf[src] * 4, ncclFloat, src, ctx->comm, ctx->gatherStream);
    }
    const int r2 = ctx->nccl.GroupEnd();
    if (r != 0 || r2 != 0)
    {
        ctx->cur = ctx->stream;
        return fail(ctx, FLX_E_NCCL, "NCCL gather failed: %s", ctx->nccl.GetErrorString(r != 0 ? r : r2));
    }
    if (isRoot && !direct)
        k_deinterleave<<<(unsigned)((fullPixels + 255) / 256), 256, 0, ctx->gatherStream>>>(reinterpret_cast<const float4 *>(ctx->gatherBuf),
                                                                                            reinterpret_cast<float4 *>(ctx->fullImage), ctx->width, ctx->height,
                                                                                            ctx->nParts, ctx->stripeRows, maxTile);
    }
    ctx->cur = ctx->stream;
    if ((rc = launchCheck(ctx, "k_deinterleave")))
        return rc;
    CU(cudaEventRecord(ctx->evGatherDone, ctx->gatherStream));
    ctx->gatherInFlight = true;
    ctx->lastGatherRoot = root;
    if (needSnapshot)
    {
        CU(cudaEventRecord(ctx->evSnapshotFree[par], ctx->gatherStream));
        ctx->snapshotBusy[par] = true;
    }
    if (ctx->rank == root && full_rgba_host_or_null)
    {
        CU(cudaMemcpyAsync(full_rgba_host_or_null, ctx->fullImage, fullPixels * 16, cudaMemcpyDeviceToHost, ctx->gatherStream));
        CU(cudaStreamSynchronize(ctx->gatherStream));
    }
    return 0;
}
FLX_API_CATCH(ctx)

// Root only, after flx_gather_pixels: the gathered full image -- the accumulators (preview = 0), or the display pass over them
// (preview != 0: mk_postprocess.cl:7-55 with the current exposure / tone-map operator, the same kernel a single-GPU context runs
// over its own image).  On the gather stream, behind the gather it reads.
int flx_read_gathered(flx_ctx *ctx, int preview, float *rgba, size_t n_pixels)
try
{
    int rc = checkReady(ctx, false, true);
    if (rc)
        return rc;
    REQUIRE(rgba != nullptr, "flx_read_gathered: null destination");
    REQUIRE(ctx->comm != nullptr && ctx->lastGatherRoot == ctx->rank && ctx->fullImage, "flx_read_gathered: this context was not the root of a gather (flx_gather_pixels)");
    const size_t fullPixels = (size_t)ctx->width * ctx->height;
    REQUIRE(n_pixels <= fullPixels && ctx->fullImagePixels >= fullPixels, "flx_read_gathered: more pixels requested than the gathered image holds");
    CU(cudaSetDevice(ctx->device));
    const float *src = ctx->fullImage;
    if (preview)
    {
        if (ctx->fullPreviewPixels < fullPixels)
        {
            CU(cudaStreamSynchronize(ctx->gatherStream));
            freeDev(ctx->fullPreview);
            ctx->fullPreviewPixels = 0;
            CU(cudaMalloc(&ctx->fullPreview, fullPixels * 16));
            ctx->fullPreviewPixels = fullPixels;
        }
        k_postprocess<<<streamingGrid((uint32_t)fullPixels), FLX_BLOCK, 0, ctx->gatherStream>>>(reinterpret_cast<const float4 *>(ctx->fullImage), reinterpret_cast<float4 *>(ctx->fullPreview),
                                                                                              nullptr, 1, (uint32_t)fullPixels, ctx->params.ppParams.exposure,
                                                                                              ctx->params.ppParams.tmOperator);
        if ((rc = launchCheck(ctx, "k_postprocess<gathered image>")))
            return rc;
        src = ctx->fullPreview;
    }
    CU(cudaMemcpyAsync(rgba, src, n_pixels * 16, cudaMemcpyDeviceToHost, ctx->gatherStream));
    CU(cudaStreamSynchronize(ctx->gatherStream));
    return 0;
}
FLX_API_CATCH(ctx)

} // extern "C"
