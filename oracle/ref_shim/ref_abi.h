/* ref_abi.h -- TEST INFRASTRUCTURE (oracle). Plain-C argument block handed to the host-compiled
 * reference kernels in oracle/_ref/libfluctus_ref.so.  Field names follow the kernel argument
 * names the reference binds in src/kernel_impl.hpp:18-48, 90-139, 246-285. */
#ifndef REF_ABI_H
#define REF_ABI_H
#include <stdint.h>
typedef struct
{
    void *tasks;            /* GPUTaskState SoA, numTasks*256 B */
    float *pixels;          /* W*H float4 */
    float *denoiserAlbedo;
    float *denoiserNormal;
    void *queueLens;        /* QueueCounters, 8 x u32 */
    uint32_t *raygenQueue, *extensionQueue, *shadowQueue;
    uint32_t *diffuseQueue, *glossyQueue, *ggxReflQueue, *ggxRefrQueue, *deltaQueue;
    void *tris;             /* Triangle[ ] 160 B */
    void *nodes;            /* GPUNode[ ] 48 B */
    uint32_t *indices;
    const float *envRGBA;   /* RGBA32F env map (or 1x1 dummy) */
    int32_t envW, envH;
    float *probTable;
    int32_t *aliasTable;
    float *pdfTable;
    void *materials;        /* Material[ ] 80 B */
    uint8_t *texData;
    void *textures;         /* TexDescriptor[ ] 12 B */
    void *params;           /* RenderParams 240 B */
    uint32_t *currPixelIdx;
    uint32_t numTasks;
    uint32_t firstIteration;
    float *pixelsPreview;   /* W*H float4, output of the post-process kernel (GL PBO in the reference) */
    uint32_t *stats;        /* RenderStats {primaryRays, extensionRays, shadowRays, samples} (geom.h:254-260), microkernel integrator */
    float *denoiserAlbedoGL; /* W*H float4 each: outputs of the post-process kernel when built with USE_OPTIX_DENOISER */
    float *denoiserNormalGL;
} RefBufs;
#endif
