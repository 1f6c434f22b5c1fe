/*
 * fluctus_b200.h -- C ABI of the B200-native wavefront path-tracing hot path.
 *
 * Drop-in boundary: the reference has no FFI; its boundary for this path is the C++ class
 * CLContext (reference: src/clcontext.hpp:26-211) as driven by Tracer::update()
 * (src/tracer.cpp:222-266) and Tracer::runBenchmark() (src/tracer.cpp:431-470), plus the
 * per-kernel argument contracts of src/kernel_impl.hpp.  Every entry point below names the
 * CLContext method it replaces.  All structs are byte-for-byte the reference's device layouts
 * (src/geom.h), so a caller can hand over the very arrays it used to write into cl::Buffers.
 *
 * Conventions
 *   - every function returns 0 on success, else a non-zero code (cudaError_t, or FLX_E_*);
 *     flx_last_error(ctx) returns the message.  Nothing ever calls exit()  (the reference
 *     throws std::runtime_error through clt::check, ext/CLT/src/utils.cpp:22-29; the C++
 *     wrapper include/fluctus_b200/clcontext.hpp rethrows to keep that behaviour).
 *   - the library owns all device memory; host pointers are borrowed for the duration of the
 *     call (uploads are blocking, like the reference's CL_TRUE writes), except the output of
 *     flx_enqueue_get_counters which is valid after the next flx_finish (reference: CL_FALSE
 *     read, src/clcontext.cpp:668-671).
 *   - one context = one GPU = one in-order CUDA stream (reference: one cl::CommandQueue);
 *     a context is not thread-safe; any number of contexts may live in one process.
 *   - there is no CPU fallback: flx_create fails when no sm_100 device is usable.
 */
#ifndef FLUCTUS_B200_H
#define FLUCTUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- layouts (reference: src/geom.h; sizes checked by static_assert in the library and by tests) */
typedef struct { float x, y, z, w; } flx_float3;  /* OpenCL float3 / FireRays::float3: 16 bytes */
typedef struct { float x, y; } flx_float2;

typedef struct { flx_float3 right, up, N, pos, E; flx_float2 size; float _pad[2]; } flx_AreaLight;      /* geom.h:104-111, 96 B */
typedef struct { flx_float3 pos, dir, up, right; float fov, apertureSize, focalDist, _pad; } flx_Camera; /* geom.h:146-155, 80 B */
typedef struct { float exposure; uint32_t tmOperator; } flx_PostProcessParams;                            /* geom.h:157-161 */
typedef struct {                                                                                          /* geom.h:163-180, 240 B */
    flx_AreaLight areaLight;
    flx_Camera camera;
    flx_PostProcessParams ppParams;
    uint32_t width, height, n_tris, useEnvMap, useAreaLight;
    float envMapStrength;
    uint32_t maxBounces, sampleImpl, sampleExpl, useRoulette, wfSeparateQueues;
    float worldRadius;
    uint32_t _pad[2];
} flx_RenderParams;

typedef struct { flx_float3 bmin, bmax; int32_t parent; uint32_t iStartOrRightChild; uint8_t nPrims; uint8_t _pad[7]; } flx_Node;  /* geom.h:71-80, bvhnode.hpp:50-59, 48 B */
typedef struct { flx_float3 p, n, t; } flx_Vertex;                                                        /* geom.h:82-87 */
typedef struct { flx_Vertex v0, v1, v2; int32_t matId; int32_t _pad[3]; } flx_Triangle;                    /* geom.h:89-95, triangle.hpp:18-49, 160 B */
typedef struct { flx_float3 Kd, Ks, Ke; float Ns, Ni; int32_t map_Kd, map_Ks, map_N, type; int32_t _pad[2]; } flx_Material; /* geom.h:113-124, 80 B */
typedef struct { uint32_t offset, width, height; } flx_TexDescriptor;                                     /* geom.h:126-131, 12 B */
typedef struct { uint32_t raygenQueue, extensionQueue, shadowQueue, diffuseQueue, glossyQueue, ggxReflQueue, ggxRefrQueue, deltaQueue; } flx_QueueCounters; /* geom.h:240-252 */
typedef struct { uint32_t primaryRays, extensionRays, shadowRays, samples; } flx_RenderStats;             /* geom.h:254-260 */
typedef struct { uint64_t primaryRays, extensionRays, shadowRays, samples, iterations; } flx_RenderStats64; /* same counters, 64-bit (new) */
typedef struct { float primary, extension, shadow, samples, total; } flx_PerfNumbers;                     /* clcontext.hpp:12-19 */
/* traversal work in the reference's terms, summed over rays (new; numerator of the roofline):
 * nodes popped (bvh.cl:251/329), child boxes tested (bvh.cl:283-284), triangles tested (bvh.cl:260),
 * closest-hit updates (bvh.cl:271-279), rays traced */
typedef struct { uint64_t nodes, boxes, tris, updates, rays; } flx_TraceCounts;

/* BSDF type bits, reference src/bxdf_types.h:4-11 */
#define FLX_BXDF_DIFFUSE (1 << 1)
#define FLX_BXDF_GLOSSY (1 << 2)
#define FLX_BXDF_GGX_ROUGH_REFLECTION (1 << 3)
#define FLX_BXDF_IDEAL_REFLECTION (1 << 4)
#define FLX_BXDF_GGX_ROUGH_DIELECTRIC (1 << 5)
#define FLX_BXDF_IDEAL_DIELECTRIC (1 << 6)
#define FLX_BXDF_EMISSIVE (1 << 7)

/* GPUTaskState is a structure of arrays: 64 four-byte slots, slot s of path g lives at
 * ((uint32_t*)tasks)[s * num_tasks + g]  (geom.h:37-49, 199-236).  Slot numbers: */
enum {
    FLX_S_ORIG = 0, FLX_S_DIR = 4, FLX_S_SHADOW_ORIG = 8, FLX_S_SHADOW_DIR = 12, FLX_S_T = 16, FLX_S_EI = 20,
    FLX_S_LAST_BSDF = 24, FLX_S_LAST_EMISSION = 28, FLX_S_LAST_T = 32, FLX_S_P = 36, FLX_S_N = 40, FLX_S_UV = 44,
    FLX_S_PHASE = 46, FLX_S_LAST_PDF_W = 47, FLX_S_PATH_LEN = 48, FLX_S_SEED = 49, FLX_S_LAST_SPECULAR = 50,
    FLX_S_SHADOW_BLOCKED = 51, FLX_S_BACKFACE = 52, FLX_S_PIXEL_INDEX = 53, FLX_S_FIRST_DIFFUSE = 54,
    FLX_S_LAST_PDF_DIRECT = 55, FLX_S_LAST_PDF_IMPLICIT = 56, FLX_S_LAST_COS_TH = 57, FLX_S_LAST_LIGHT_PICK = 58,
    FLX_S_SHADOW_RAY_LEN = 59, FLX_S_HIT_T = 60, FLX_S_HIT_I = 61, FLX_S_AREA_LIGHT_HIT = 62, FLX_S_MAT_ID = 63,
    FLX_NUM_SLOTS = 64
};

enum { FLX_OK = 0, FLX_E_INVALID = 10001, FLX_E_NO_DEVICE = 10002, FLX_E_NOT_READY = 10003, FLX_E_NCCL = 10004, FLX_E_UNSUPPORTED_ARCH = 10005 };

/* kernel ids for flx_get_kernel_ms (reference instrument: CLContext::checkTracingPerf, clcontext.cpp:673-701) */
enum { FLX_K_RESET = 0, FLX_K_RAYGEN, FLX_K_EXTRAYS, FLX_K_SHADOWRAYS, FLX_K_LOGIC, FLX_K_MATERIALS, FLX_K_END_ITERATION, FLX_K_POSTPROCESS,
       FLX_K_MK_RESET, FLX_K_MK_RAYGEN, FLX_K_MK_NEXT_VERTEX, FLX_K_MK_SAMPLE_BSDF, FLX_K_MK_SPLAT,
       FLX_K_LOGIC_FUSED /* logic + raygen + materials in one kernel (flx_render) */,
       FLX_K_GATHER /* the NCCL gather of the tile accumulators + de-interleave, on its own stream (flx_gather_pixels) */, FLX_K_COUNT };

typedef struct flx_ctx flx_ctx;

/* CLContext::CLContext + setup (clcontext.hpp:32,62; clcontext.cpp:18-69,116-141): pick the device, allocate the
 * path-state SoA for num_tasks paths (256 B each), 8 index queues, counters, pixel index, params. */
int flx_create(int device, uint32_t num_tasks, flx_ctx **out);
void flx_destroy(flx_ctx *ctx);
const char *flx_last_error(const flx_ctx *ctx); /* ctx may be NULL: message of the last failed flx_create */
const char *flx_version(void);

/* CLContext::uploadSceneData + packTextures (clcontext.hpp:76; clcontext.cpp:522-611). Blocking; copies.
 * nodes: reference Node[] in DFS order, left child = self+1 (bvhnode.hpp:50-59); indices: u32 triangle refs. */
int flx_upload_scene(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, const uint32_t *indices, uint32_t n_indices,
                     const flx_Node *nodes, uint32_t n_nodes, const flx_Material *materials, uint32_t n_materials,
                     const flx_TexDescriptor *tex_desc, uint32_t n_tex, const uint8_t *tex_data, size_t tex_bytes);

/* GPU hierarchy builder (new).  Stands in for the reference's CPU builders -- `new SBVH(&tris, mode)` /
 * `new BVH(...)` in Scene's initHierarchy (src/scene.cpp:574-590; src/sbvh.cpp:4-449, src/bvh.cpp:205-407) -- and returns THE
 * SAME two arrays they produce (BVH::m_nodes, m_indices): Node[] in depth-first order with left child = self + 1
 * (src/bvhnode.hpp:50-59, src/sbvh.cpp:52-73) and the u32 triangle index list, ready for flx_upload_scene.  LBVH (Morton
 * order + binary radix tree) with a bottom-up SAH collapse into leaves of at most max_leaf triangles (reference MaxLeafElems
 * = 8, src/bvh.hpp:70); no spatial splits, so every triangle is referenced exactly once (n_indices = n_tris).  Milliseconds
 * instead of seconds; the tree is of lower quality than the reference's SBVH (DESIGN.md 4.5 has the measured trade-off).
 * nodes_out must hold nodes_capacity >= 2 * n_tris - 1 records in the worst case; build_ms (may be NULL) = device time. */
enum { FLX_BVH_FAST = 0, /* LBVH: Morton order + binary radix tree + SAH collapse; a third of a millisecond for 300 k triangles */
       FLX_BVH_PLOC = 1, /* parallel locally-ordered clustering (radius 16) on the Morton order + SAH collapse: better trees, a few ms */
       FLX_BVH_PLOC_OPT = 2 /* FLX_BVH_PLOC + parallel reinsertion (FLX_TUNE_BVH_REINSERT iterations, default 16): subtrees move to where they
                               enlarge the fewest boxes -- the overlap reduction the reference's SBVH gets from spatial splits (src/sbvh.cpp:118-142);
                               trees trace within 3 % of / up to 5 % faster than the reference's SBVH; tens of ms */ };
int flx_build_bvh(flx_ctx *ctx, const flx_Triangle *tris, uint32_t n_tris, uint32_t max_leaf, int quality, flx_Node *nodes_out, uint32_t nodes_capacity,
                  uint32_t *n_nodes_out, uint32_t *indices_out /* n_tris */, float *build_ms);

/* CLContext::createEnvMap (clcontext.hpp:79; clcontext.cpp:467-511): rgb is w*h*3 floats; tables are w*h entries. */
int flx_upload_envmap(flx_ctx *ctx, const float *rgb, int32_t w, int32_t h, const float *prob, const int32_t *alias, const float *pdf);

/* CLContext::setupPixelStorage (clcontext.hpp:77; clcontext.cpp:326-384) without the GL PBOs: (re)allocate W*H float4. */
int flx_resize(flx_ctx *ctx, uint32_t width, uint32_t height);

/* CLContext::updateParams (clcontext.hpp:75; clcontext.cpp:703-707); also plays the role of recompileKernels
 * (clcontext.cpp:852-874): kernel specialisations are picked from the params at launch time. */
int flx_update_params(flx_ctx *ctx, const flx_RenderParams *params);

/* CLContext::enqueueWf*Kernel (clcontext.hpp:43-48; clcontext.cpp:765-848). Asynchronous, in order. */
int flx_enqueue_reset(flx_ctx *ctx);
int flx_enqueue_raygen(flx_ctx *ctx);
int flx_enqueue_extrays(flx_ctx *ctx);
int flx_enqueue_shadowrays(flx_ctx *ctx);
int flx_enqueue_logic(flx_ctx *ctx, int first_iteration);
int flx_enqueue_materials(flx_ctx *ctx); /* 5 per-type kernels or the single-queue kernel, by params.wfSeparateQueues */

/* CLContext::enqueuePostprocessKernel (clcontext.hpp:41; clcontext.cpp:752-763; kernel src/mk_postprocess.cl:7-55): the display
 * pass of every loop iteration -- divide by the sample count, exposure, tone map (params.ppParams), gamma -- into a float4
 * preview buffer (a GL PBO in the reference; here read back with flx_read_preview). */
int flx_enqueue_postprocess(flx_ctx *ctx);
int flx_read_preview(flx_ctx *ctx, float *rgba, size_t n_pixels);

/* Denoiser feature buffers (reference: Tracer::useDenoiser -> -DUSE_OPTIX_DENOISER on the logic, nextVertex, sampleBsdf and
 * post-process kernels, src/kernel_impl.hpp:53,346,380,443; the code is src/wf_logic.cl:186-209, src/mk_next_vertex.cl:60-70,
 * src/mk_sample_bsdf.cl:56-66, src/mk_postprocess.cl:49-54).  While enabled, both integrators accumulate the first-hit shading
 * normal in camera space and the albedo at the path's first non-singular vertex per pixel (RGB sums, w = sample count), and the
 * display pass also writes both divided by their count.  The OptiX denoiser itself is not part of this library: attach any.
 * flx_read_denoiser_aov: which 0 = normal, 1 = albedo; processed 0 = raw accumulator, 1 = display-pass output. */
int flx_set_denoiser(flx_ctx *ctx, int enabled);
int flx_read_denoiser_aov(flx_ctx *ctx, int which, int processed, float *rgba, size_t n_pixels);

/* The reference's other integrator, the "microkernel" path tracer (one path per pixel, a phase word per path): the one
 * Tracer::renderSingle uses for final frames with an exact sample count per pixel (src/tracer.cpp:95-169) and the
 * non-wavefront branch of Tracer::update (src/tracer.cpp:267-299).  CLContext::enqueueResetKernel / enqueueRayGenKernel /
 * enqueueNextVertexKernel / enqueueBsdfSampleKernel / enqueueSplatKernel / enqueueSplatPreviewKernel (clcontext.hpp:34-40;
 * clcontext.cpp:709-750; kernels src/mk_reset.cl, mk_raygen.cl, mk_next_vertex.cl, mk_sample_bsdf.cl, mk_splat.cl,
 * mk_splat_preview.cl).  Paths 0 .. min(width*height, num_tasks)-1 take part, path g renders pixel g; ray and sample counts
 * accumulate in the 64-bit stats (reference: atomics on RenderStats).  The path state after every call is the reference's. */
int flx_enqueue_mk_reset(flx_ctx *ctx);
int flx_enqueue_mk_raygen(flx_ctx *ctx);
int flx_enqueue_mk_next_vertex(flx_ctx *ctx);
int flx_enqueue_mk_sample_bsdf(flx_ctx *ctx);
int flx_enqueue_mk_splat(flx_ctx *ctx);
int flx_enqueue_mk_splat_preview(flx_ctx *ctx);
/* The sample loop of Tracer::renderSingle (src/tracer.cpp:124-150) spp times without host round trips: camera rays,
 * (maxBounces + 1) x (nextVertex, sampleBsdf), splat, display pass.  Call flx_enqueue_mk_reset first, like the reference. */
int flx_render_single(flx_ctx *ctx, uint32_t spp);

/* CLContext::enqueueClearWfQueues / enqueueGetCounters / finishQueue / updatePixelIndex / resetPixelIndex /
 * getNumTasks (clcontext.hpp:53-57,71; clcontext.cpp:668-671, 877-906). */
int flx_enqueue_clear_queues(flx_ctx *ctx);
int flx_enqueue_get_counters(flx_ctx *ctx, flx_QueueCounters *host_out);
int flx_finish(flx_ctx *ctx);
int flx_update_pixel_index(flx_ctx *ctx, uint32_t num_pixels, uint32_t num_new_paths);
int flx_reset_pixel_index(flx_ctx *ctx);
uint32_t flx_num_tasks(const flx_ctx *ctx);

/* The steady-state loop of Tracer::runBenchmark (tracer.cpp:431-470) replayed n_iterations times with no host
 * round trip: logic, raygen, materials, [counter snapshot], extrays, shadowrays, then a one-thread kernel that does
 * what the host does between iterations (stats += counters, pixelIdx = (pixelIdx + cnt.raygen) % numPixels
 * [clcontext.cpp:891-895], counters = 0).  Results are identical to driving the single calls above. */
int flx_render(flx_ctx *ctx, uint32_t n_iterations);

/* flx_render bracketed by CUDA events on the context's stream; elapsed_ms is device time (synchronises). */
int flx_render_timed(flx_ctx *ctx, uint32_t n_iterations, float *elapsed_ms);
/* a CUDA-event pair on the context's stream around any sequence of calls (begin synchronises first, end waits) */
int flx_timer_begin(flx_ctx *ctx);
int flx_timer_end(flx_ctx *ctx, float *elapsed_ms);

/* Tuning knobs; results never depend on them (tests/test_gpu_parity.py runs the parity suite over the variants). */
enum { FLX_TUNE_TRACE_VARIANT = 0,      /* 0: one ray per thread; 1 (default): persistent threads with dynamic ray fetch, while-while phases; 2: 1 + the
                                           top-of-tree treelet staged in shared memory by the bulk-copy engine, one CTA per SM; 3: persistent threads, one
                                           step of the majority kind (inner node / one triangle) per iteration (flx_trace_greedy.cuh; measured equal to 1) */
       FLX_TUNE_BVH_TRI_COST = 22,      /* flx_build_bvh: cost of a triangle test relative to a box test in the SAH collapse decision, in percent
                                           (100 = the reference's constants costTri = costBox = 1, src/bvh.hpp:72-73; larger = smaller leaves) */
       FLX_TUNE_MATERIAL_MASK = 27,     /* fused logic kernel: 1 (default) compile in only the BSDF lobes the uploaded materials use, like the reference's kernel
                                           build (src/kernel_impl.hpp:261-266); 0 always the all-lobes instantiation */
       FLX_TUNE_BVH_DEPTH_LIMIT = 28,   /* flx_build_bvh: deepest PLOC tree handed out, 1..62 (default 62: the traversal stack holds 64 entries).  A PLOC_OPT tree
                                           beyond it falls back to the PLOC tree it started from; a PLOC tree beyond it is an error */
       FLX_TUNE_BVH_REINSERT = 26,      /* flx_build_bvh(FLX_BVH_PLOC_OPT): iterations of the reinsertion post-pass, 0..64 (default 16) */
       FLX_TUNE_SHADOW_LEFT_FIRST = 25, /* (removed: any-hit traversal taking the left child first was measured slower; only 0 is accepted) */
       FLX_TUNE_LOGIC_TILE = 24,        /* paths per tile (= threads per CTA) of the logic kernel: 256 (default) or 128 */
       FLX_TUNE_GATHER_DIRECT = 23,     /* flx_gather_pixels: 0 (default) one send / receive per rank into a rank-major buffer + a de-interleave pass; 1 one per
                                           stripe, straight into the rows of the root's full image (measured 3-4x slower: NCCL's per-operation cost); must
                                           be the same on every rank */
       FLX_TUNE_GATHER_PRIORITY = 21,   /* stream priority of the NCCL gather: 0 (default) lowest, 1 = the render stream's; set before the first gather */
       FLX_TUNE_INNER_BIAS = 20,        /* variant 3: run an inner-node step when lanes-at-inner + bias >= lanes-at-a-triangle (default 0) */
       FLX_TUNE_FETCH_THRESHOLD = 1,    /* persistent variant: refill a warp when fewer lanes than this hold a ray (default 16) */
       FLX_TUNE_TRACE_BLOCKS_PER_SM = 2,/* variant 1: resident CTAs per SM, 0 = occupancy calculator */
       FLX_TUNE_TOP_NODES = 3,          /* variant 2: treelet nodes (64 B each) staged per CTA, default 2047 */
       FLX_TUNE_LOGIC_MIN_BLOCKS = 5,   /* register budget of the logic kernel: compiled for 2, 3 (default) or 4 resident CTAs per SM */
       FLX_TUNE_FETCH_CHUNK = 6,        /* persistent variants: queue entries a warp reserves per atomic (default 32) */
       FLX_TUNE_EXT_MIN_BLOCKS = 8,     /* variant 1 register budget: extension kernel compiled for 8, 9 (default) or 10 CTAs of 128 per SM */
       FLX_TUNE_SHADOW_MIN_BLOCKS = 9,  /* same for the shadow kernel (default 10) */
       FLX_TUNE_POSTPROCESS_IN_LOOP = 10, /* flx_render: run the display pass every iteration like the reference's loop (default 1) */
       FLX_TUNE_SMEM_STACK = 11,        /* variant 1: keep the first 4 / 8 / 24 (value; 1 = 24) traversal-stack levels in shared memory; 0 = local memory; -1 = local memory + newest entry in a register */
       FLX_TUNE_MAX_L1 = 12,            /* variant 1: request the maximum L1 carve-out for the traversal kernels (default 0: measured 4 % slower) */
       FLX_TUNE_OVERLAP_TRACE = 7,      /* flx_render: run the shadow-ray kernel on a second stream, overlapping the extension kernel's tail (default 1) */
       FLX_TUNE_L2_PERSIST = 19,          /* persisting-L2 access window for the traversal streams: 0 off (default), 1 over the TTri array, 2 over the TNode array;
                                             set after flx_upload_scene */
       FLX_TUNE_DIRTY_POSTPROCESS = 18,   /* display pass recomputes only the pixels whose accumulator changed since its last run (default 1) */
       FLX_TUNE_OVERLAP_POSTPROCESS = 17, /* display pass beside the traversal stages on a third stream: 0 never, 1 (default) in the per-stage ABI
                                             (flx_enqueue_postprocess starts from the accumulator's last writer), 2 also inside flx_render */
       FLX_TUNE_REPACK_ON_HOST = 16,    /* flx_upload_scene: make the traversal layout with the host code instead of the device kernels (default 0) */
       FLX_TUNE_PREFETCH_CHILDREN = 15, /* (removed: prefetching both children of an inner node was measured slower; only 0 is accepted) */
       FLX_TUNE_FUSE_STAGES = 13,       /* flx_render: logic + raygen + materials as ONE kernel over the path state (default 1) */
       FLX_TUNE_FUSED_MIN_BLOCKS = 14,  /* register budget of that kernel: compiled for 1..4 resident CTAs of 256 per SM (default 3) */
       FLX_TUNE_INNER_MIN = 4           /* leave the inner-node phase when fewer lanes than this are still at inner nodes (default 8) */ };
int flx_set_tuning(flx_ctx *ctx, int key, int value);

/* Instrumented traversal: while enabled, flx_enqueue_extrays / flx_enqueue_shadowrays also count the work the
 * reference's algorithm does per ray (results are unchanged). Used outside timed regions only. */
int flx_set_counting(flx_ctx *ctx, int enabled);
int flx_get_trace_counts(flx_ctx *ctx, flx_TraceCounts *ext, flx_TraceCounts *shadow);

/* CLContext::resetStats/getStats/updateRenderPerf/getRenderPerf (clcontext.hpp:66-70; clcontext.cpp:634-666).
 * Totals are accumulated on the device by flx_render (64-bit) and are read here (synchronises). */
int flx_reset_stats(flx_ctx *ctx);
int flx_get_stats(flx_ctx *ctx, flx_RenderStats64 *out);

/* CLContext::checkTracingPerf (clcontext.hpp:73; clcontext.cpp:673-701): accumulated CUDA-event time and launch
 * count of one kernel since the last flx_reset_stats. Timing is off by default (flx_set_profiling). */
int flx_set_profiling(flx_ctx *ctx, int enabled);
int flx_get_kernel_ms(flx_ctx *ctx, int kernel_id, float *total_ms, uint32_t *launches);

/* Raw read-back, replaces CLContext::saveImage (clcontext.hpp:78; clcontext.cpp:386-465). Synchronises. */
int flx_read_pixels(flx_ctx *ctx, float *rgba, size_t n_pixels);
/* CLContext::saveImage itself (clcontext.cpp:386-465): "*.hdr" = accumulator / sample count, linear, as Radiance RGBE; any
 * other name = the post-processed preview (run flx_enqueue_postprocess first) as 8-bit PNG, byte = (uchar)(255 * clamp01(c)).
 * Row 0 of the buffers is the bottom image row (the reference sets DevIL's origin to lower-left, src/main.cpp:69-71); files
 * are written top row first.  flx_write_image does the same conversions on a caller's RGBA float buffer (host code). */
int flx_save_image(flx_ctx *ctx, const char *filename);
int flx_write_image(const char *path, const float *rgba, uint32_t width, uint32_t height);

/* Checkpoint / resume (new).  The reference can only restart a render: its caches hold the hierarchy, camera state and kernel
 * binaries, never the accumulator.  One file = path state, queues, counters, pixel index, statistics, accumulator of
 * this context; an interrupted wavefront or microkernel render continues from it with a bit-identical path state.  Scene,
 * environment map, image size, tile and params are not stored: set them up as usual (same num_tasks), then load. */
int flx_checkpoint_save(flx_ctx *ctx, const char *path);
int flx_checkpoint_load(flx_ctx *ctx, const char *path);

/* Test/diagnostic access to the path state and queues (no reference equivalent; the reference's debugger did this). */
int flx_read_tasks(flx_ctx *ctx, uint32_t *slots_out /* 64*num_tasks */);
/* the hierarchy as repacked for traversal (16 floats per inner node / per leaf reference); NULL arrays: counts only */
int flx_read_traversal_layout(flx_ctx *ctx, float *tnodes_out, uint32_t *n_tnodes, float *ttris_out, uint32_t *n_ttris, int32_t *root_ref);
int flx_write_tasks(flx_ctx *ctx, const uint32_t *slots_in);
int flx_read_queue(flx_ctx *ctx, int queue_id /* 0..7 = order of flx_QueueCounters */, uint32_t *out, uint32_t max_entries);
int flx_write_queue(flx_ctx *ctx, int queue_id, const uint32_t *entries, uint32_t n);
int flx_write_counters(flx_ctx *ctx, const flx_QueueCounters *in);

/* ---- image-space sharding across GPUs (new). A context renders the rows {y : (y / stripe_rows) % n_parts
 * == part} of a full_width x full_height image; its local pixel buffer holds only those rows, top to bottom.
 * flx_set_tile must be followed by flx_update_params (width/height there are the FULL image). */
int flx_set_tile(flx_ctx *ctx, uint32_t part, uint32_t n_parts, uint32_t stripe_rows);
uint32_t flx_tile_pixels(const flx_ctx *ctx);

/* One NCCL collective per frame: gather every rank's tile into the full image on `root` (de-interleaved on the GPU).
 * NCCL is resolved at run time (dlopen of libnccl.so.2, i.e. the copy torch.distributed already loaded). */
int flx_comm_unique_id(void *out128);
int flx_comm_init(flx_ctx *ctx, const void *unique_id128, int rank, int nranks);
/* Asynchronous like the enqueue calls: the frame gathered is the accumulator as of this call (a device-side snapshot taken in
 * stream order), the transfer runs on the library's gather stream beside whatever is enqueued next, and is complete after
 * flx_finish -- or on return when a host destination is given.  Device time per call: flx_get_kernel_ms(FLX_K_GATHER). */
int flx_gather_pixels(flx_ctx *ctx, int root, float *full_rgba_host_or_null);
/* Root only, after flx_gather_pixels: the gathered full image read back -- the accumulators (preview = 0) or the display pass over
 * them (preview != 0: the post-process kernel of the reference, mk_postprocess.cl:7-55, with the current exposure / tone-map
 * operator).  flx_save_image on the root of a tiled context writes this frame (.hdr: accumulators, otherwise the display pass). */
int flx_read_gathered(flx_ctx *ctx, int preview, float *rgba, size_t n_pixels);
int flx_comm_destroy(flx_ctx *ctx);

/* ---- scene input (host code, no device work).  The reference's Scene::loadModel for OBJ + MTL and ASCII PLY
 * (src/scene.cpp:52-92, 191-301, 422-553 with its vendored tinyobjloader 1.0.x) and its EnvironmentMap (src/envmap.cpp:9-114
 * with src/rgbe/rgbe.cpp), producing exactly the arrays flx_upload_scene / flx_upload_envmap take: triangles in file order,
 * material 0 = the reference's default material, texture NAMES in first-use order (decoding images is left to the caller;
 * the reference uses DevIL), RGB float image + alias-method tables.  Errors: non-zero return, message in flx_io_last_error(). */
typedef struct flx_scene flx_scene;
typedef struct flx_envmap flx_envmap;
const char *flx_io_last_error(void);
int flx_scene_load(const char *path /* .obj or .ply */, flx_scene **out);
void flx_scene_free(flx_scene *scene);
uint32_t flx_scene_num_triangles(const flx_scene *scene);
uint32_t flx_scene_num_materials(const flx_scene *scene);
uint32_t flx_scene_num_textures(const flx_scene *scene);
const flx_Triangle *flx_scene_triangles(const flx_scene *scene);
const flx_Material *flx_scene_materials(const flx_scene *scene);
const char *flx_scene_texture_name(const flx_scene *scene, uint32_t index); /* relative to the model's folder */
int flx_envmap_load(const char *path /* Radiance .hdr */, flx_envmap **out);
int flx_envmap_from_rgb(const float *rgb, int32_t w, int32_t h, flx_envmap **out); /* tables only (EnvironmentMap::computeProbabilities) */
void flx_envmap_free(flx_envmap *env);
int32_t flx_envmap_width(const flx_envmap *env);
int32_t flx_envmap_height(const flx_envmap *env);
const float *flx_envmap_rgb(const flx_envmap *env);
const float *flx_envmap_prob(const flx_envmap *env);
const int32_t *flx_envmap_alias(const flx_envmap *env);
const float *flx_envmap_pdf(const flx_envmap *env);

/* Textures.  flx_image_load decodes a PNG to the reference's in-memory form (RGBA8, row 0 = bottom row; src/texture.cpp:16-40 with
 * DevIL's origin at lower-left, src/main.cpp:69-71); JPEG is not decoded (lossy, decoder-dependent: SURVEY 8c) -- pass such images
 * decoded.  flx_pack_textures = CLContext::packTextures (src/clcontext.cpp:570-611): descriptors + images back to back, the two
 * arrays flx_upload_scene takes; with blob_out == NULL it only reports the size. */
int flx_image_load(const char *path, uint32_t *width, uint32_t *height, uint8_t **rgba);
void flx_image_free(uint8_t *rgba);
int flx_pack_textures(const uint8_t *const *images, const uint32_t *widths, const uint32_t *heights, uint32_t n_tex, flx_TexDescriptor *desc_out, uint8_t *blob_out,
                      size_t *blob_bytes);

/* The reference's hierarchy cache file (BVH::exportTo / importFrom, src/bvh.cpp:102-192; `data/hierarchies/hierarchy_<hash>.bin`,
 * src/tracer.cpp:574-590, 742-751), both directions, so caches can be exchanged with the reference.  The reference writes the
 * index count where the node count belongs (src/bvh.cpp:185): the importer here derives the node count from the file length,
 * the exporter writes the true count (which the reference's importer handles).  Import is a two-call protocol: pass NULL arrays
 * to get the counts, then arrays of at least that size (n_nodes / n_indices in: capacity, out: count). */
int flx_hierarchy_export(const char *path, const flx_Node *nodes, uint32_t n_nodes, const uint32_t *indices, uint32_t n_indices);
int flx_hierarchy_import(const char *path, flx_Node *nodes_out, uint32_t *n_nodes, uint32_t *indices_out, uint32_t *n_indices);

/* Page-locked host memory (new).  Arrays handed to flx_upload_scene / flx_upload_envmap or filled by flx_read_pixels /
 * flx_read_preview / flx_gather_pixels may live anywhere; when they live in memory from flx_host_alloc the copy is a DMA at link
 * speed instead of a trip through the driver's staging buffer.  The reference's analogue is the GL pixel-buffer object its
 * kernels render into (src/clcontext.cpp:326-384).  Error text: flx_last_error(NULL). */
int flx_host_alloc(void **out, size_t bytes);
void flx_host_free(void *p);

/* device memory held by the context: path state + queues + scene + image buffers */
size_t flx_device_bytes(const flx_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FLUCTUS_B200_H */
