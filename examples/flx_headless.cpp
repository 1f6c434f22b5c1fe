// flx_headless.cpp -- headless C++ driver over the CLContext wrapper: the reference's benchmark loop
// (Tracer::runBenchmark, src/tracer.cpp:431-470) with the call sites as they appear there.
//
//   flx_headless <scene.bin> <width> <height> <numTasks> <maxBounces> <iterations> [out.rgba]
//
// <scene.bin> is a scene blob written by oracle/ref_shim/scene_tool.cpp (the reference's loader + SBVH builder); camera and
// light are the Conference set-up of SURVEY 8(d) unless the blob is not conference, in which case the reference's default
// camera/light (src/tracer.cpp:760-797) are used.  Prints Mrays/s; optionally writes the raw RGBA32F accumulator.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

#include "fluctus_b200/clcontext.hpp"

using namespace fluctus_b200;

static flx_float3 f3(float x, float y, float z, float w = 0.0f) { return flx_float3{x, y, z, w}; }
static flx_float3 normalize(flx_float3 v)
{
    const float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return f3(v.x / l, v.y / l, v.z / l);
}
static flx_float3 cross(flx_float3 a, flx_float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

int main(int argc, char **argv)
{
    if (argc < 7)
    {
        std::fprintf(stderr, "usage: %s scene.bin width height numTasks maxBounces iterations|spp=N [out.rgba]\n"
                             "  iterations: wavefront loop (Tracer::runBenchmark);  spp=N: final frame with the microkernel integrator (Tracer::renderSingle)\n", argv[0]);
        return 2;
    }
    const std::string path = argv[1];
    const bool single = std::strncmp(argv[6], "spp=", 4) == 0;
    const uint32_t W = std::atoi(argv[2]), H = std::atoi(argv[3]), N = std::atoi(argv[4]), bounces = std::atoi(argv[5]), iters = std::atoi(single ? argv[6] + 4 : argv[6]);
    std::ifstream in(path, std::ios::binary);
    if (!in)
    {
        std::fprintf(stderr, "cannot open %s\n", path.c_str());
        return 1;
    }
    uint32_t hdr[6];
    in.read(reinterpret_cast<char *>(hdr), sizeof hdr);
    if (hdr[0] != 0x53584c46u)
    {
        std::fprintf(stderr, "%s is not a scene blob\n", path.c_str());
        return 1;
    }
    std::vector<flx_Triangle> tris(hdr[1]);
    std::vector<uint32_t> indices(hdr[2]);
    std::vector<flx_Node> nodes(hdr[3]);
    std::vector<flx_Material> mats(hdr[4]);
    in.read(reinterpret_cast<char *>(tris.data()), tris.size() * sizeof(flx_Triangle));
    in.read(reinterpret_cast<char *>(indices.data()), indices.size() * 4);
    in.read(reinterpret_cast<char *>(nodes.data()), nodes.size() * sizeof(flx_Node));
    in.read(reinterpret_cast<char *>(mats.data()), mats.size() * sizeof(flx_Material));
    if (hdr[5] != 0)
        for (auto &m : mats) // this driver does not decode images: drop texture references
            m.map_Kd = m.map_Ks = m.map_N = -1;

    RenderParams params;
    std::memset(&params, 0, sizeof params);
    const bool conference = path.find("conference") != std::string::npos;
    const flx_float3 pos = conference ? f3(-0.80f, 0.05f, 0.50f) : f3(0.0f, 1.0f, 3.5f);
    const flx_float3 target = conference ? f3(0.60f, -0.08f, -0.30f) : f3(0.0f, 1.0f, 2.5f);
    params.camera.pos = pos;
    params.camera.dir = normalize(f3(target.x - pos.x, target.y - pos.y, target.z - pos.z));
    params.camera.right = normalize(cross(params.camera.dir, f3(0, 1, 0)));
    params.camera.up = cross(params.camera.right, params.camera.dir);
    params.camera.fov = 60.0f;
    params.camera.apertureSize = 0.0f;
    params.camera.focalDist = 0.5f;
    if (conference)
    {
        params.areaLight.pos = f3(0.0f, 0.235f, 0.0f);
        params.areaLight.N = f3(0.0f, -1.0f, 0.0f);
        params.areaLight.right = f3(1.0f, 0.0f, 0.0f);
        params.areaLight.up = f3(0.0f, 0.0f, 1.0f);
        params.areaLight.size = flx_float2{0.25f, 0.25f};
    }
    else
    {
        params.areaLight.pos = f3(1.0f, 1.0f, 0.0f, 1.0f);
        params.areaLight.N = f3(-1.0f, 0.0f, 0.0f);
        params.areaLight.right = f3(0.0f, 0.0f, -1.0f);
        params.areaLight.up = f3(0.0f, 1.0f, 0.0f);
        params.areaLight.size = flx_float2{0.5f, 0.5f};
    }
    params.areaLight.E = f3(200.0f, 200.0f, 200.0f);
    params.ppParams.exposure = 1.0f;
    params.ppParams.tmOperator = 2;
    params.width = W;
    params.height = H;
    params.n_tris = (uint32_t)tris.size();
    params.useAreaLight = 1;
    params.envMapStrength = 1.0f;
    params.maxBounces = bounces;
    params.sampleImpl = params.sampleExpl = 1;
    const flx_float3 d = f3(nodes[0].bmax.x - nodes[0].bmin.x, nodes[0].bmax.y - nodes[0].bmin.y, nodes[0].bmax.z - nodes[0].bmin.z);
    params.worldRadius = std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z) * 0.5f; // src/tracer.cpp:66-67

    try
    {
        CLContext clctx(N);
        SceneArrays s;
        s.tris = tris.data(); s.numTris = (uint32_t)tris.size();
        s.indices = indices.data(); s.numIndices = (uint32_t)indices.size();
        s.nodes = nodes.data(); s.numNodes = (uint32_t)nodes.size();
        s.materials = mats.data(); s.numMaterials = (uint32_t)mats.size();
        clctx.uploadSceneData(s);
        clctx.setupPixelStorage(W, H);
        clctx.updateParams(params);

        if (single) // Tracer::renderSingle (src/tracer.cpp:95-169): exactly `iters` samples in every pixel, microkernel integrator
        {
            clctx.enqueueResetKernel(params);
            const auto t0 = std::chrono::steady_clock::now();
            for (uint32_t sample = 0; sample < iters; sample++)
            {
                clctx.enqueueRayGenKernel(params);
                for (uint32_t bounce = 0; bounce < params.maxBounces + 1; bounce++)
                {
                    clctx.enqueueNextVertexKernel(params);
                    clctx.enqueueBsdfSampleKernel(params);
                }
                clctx.enqueueSplatKernel(params);
                clctx.enqueuePostprocessKernel(params);
                clctx.finishQueue();
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            flx_RenderStats64 st;
            if (flx_get_stats(clctx.handle(), &st) != 0)
                throw std::runtime_error(flx_last_error(clctx.handle()));
            std::vector<float> pix = clctx.readPixels();
            double sum[4] = {0, 0, 0, 0};
            for (size_t i = 0; i < pix.size(); i += 4)
                for (int c = 0; c < 4; c++)
                    sum[c] += pix[i + c];
            std::printf("{\"integrator\": \"microkernel\", \"spp\": %u, \"primary\": %llu, \"extension\": %llu, \"shadow\": %llu, \"seconds\": %.6f, \"mrays_per_s\": %.2f, \"samples\": %.0f, "
                        "\"mean_rgb\": [%.6f, %.6f, %.6f]}\n",
                        iters, (unsigned long long)st.primaryRays, (unsigned long long)st.extensionRays, (unsigned long long)st.shadowRays, dt,
                        (st.primaryRays + st.extensionRays + st.shadowRays) / dt / 1e6, sum[3], sum[0] / sum[3], sum[1] / sum[3], sum[2] / sum[3]);
            if (argc > 7)
            {
                std::ofstream out(argv[7], std::ios::binary);
                out.write(reinterpret_cast<const char *>(pix.data()), pix.size() * sizeof(float));
            }
            return 0;
        }

        // iteration == 0 prologue (src/tracer.cpp:236-240)
        clctx.resetPixelIndex();
        clctx.enqueueWfResetKernel(params);
        clctx.enqueueWfRaygenKernel(params);
        clctx.enqueueWfExtRayKernel(params);
        clctx.enqueueClearWfQueues();
        clctx.finishQueue();

        uint64_t ext = 0, shadow = 0, primary = 0;
        const auto t0 = std::chrono::steady_clock::now();
        for (uint32_t i = 0; i < iters; i++) // src/tracer.cpp:431-465
        {
            QueueCounters cnt;
            clctx.enqueueWfLogicKernel(params, false);
            clctx.enqueueWfRaygenKernel(params);
            clctx.enqueueWfMaterialKernels(params);
            clctx.enqueueGetCounters(&cnt);
            clctx.enqueueWfExtRayKernel(params);
            clctx.enqueueWfShadowRayKernel(params);
            clctx.enqueueClearWfQueues();
            clctx.enqueuePostprocessKernel(params);
            clctx.finishQueue();
            ext += cnt.extensionQueue;
            shadow += cnt.shadowQueue;
            primary += cnt.raygenQueue;
            clctx.updatePixelIndex(params.width * params.height, cnt.raygenQueue);
        }
        const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        std::vector<float> pix = clctx.readPixels();
        double sum[4] = {0, 0, 0, 0};
        for (size_t i = 0; i < pix.size(); i += 4)
            for (int c = 0; c < 4; c++)
                sum[c] += pix[i + c];
        std::printf("{\"iterations\": %u, \"primary\": %llu, \"extension\": %llu, \"shadow\": %llu, \"seconds\": %.6f, \"mrays_per_s\": %.2f, \"samples\": %.0f, \"mean_rgb\": [%.6f, %.6f, %.6f]}\n",
                    iters, (unsigned long long)primary, (unsigned long long)ext, (unsigned long long)shadow, dt, (ext + shadow) / dt / 1e6, sum[3], sum[0] / sum[3],
                    sum[1] / sum[3], sum[2] / sum[3]);
        if (argc > 7)
        {
            std::ofstream out(argv[7], std::ios::binary);
            out.write(reinterpret_cast<const char *>(pix.data()), pix.size() * sizeof(float));
        }
    }
    catch (const std::runtime_error &e)
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
