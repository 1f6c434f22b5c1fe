// flx_render_file.cpp -- from a model file to a picture without any of the reference's code: load OBJ/PLY through the C ABI
// (flx_scene_load), build the hierarchy on the GPU (flx_build_bvh), upload, render a final frame the way
// Tracer::renderSingle does (src/tracer.cpp:95-169: microkernel integrator, exact sample count per pixel), run the display
// pass and save it (CLContext::saveImage, src/clcontext.cpp:386-465).
//
//   flx_render_file model.obj|ply out.png|hdr [width height spp maxBounces [envmap.hdr strength]]
//
// Camera: there are no saved camera states in the reference tree (data/states is empty), so the camera is placed to frame
// the scene's bounding box; the light is the reference's "headlamp" (Tracer::updateAreaLight, src/tracer.cpp:820-826:
// the area light sits just behind the camera and faces along the view direction).  PNG and JPEG textures are decoded by the
// library (flx_image_load) and packed like CLContext::packTextures; a material whose texture file is missing or in another
// format falls back to its constants.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <fluctus_b200/clcontext.hpp>

using namespace fluctus_b200;

static flx_float3 f3(float x, float y, float z) { return flx_float3{x, y, z, 0.0f}; }
static flx_float3 norm(flx_float3 v)
{
    const float l = std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return f3(v.x / l, v.y / l, v.z / l);
}
static flx_float3 cross(flx_float3 a, flx_float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }

int main(int argc, char **argv)
{
    if (argc < 3)
    {
        std::fprintf(stderr, "usage: %s model.obj|ply out.png|hdr [width height spp maxBounces [envmap.hdr strength]]\n", argv[0]);
        return 2;
    }
    const uint32_t W = argc > 3 ? std::atoi(argv[3]) : 640, H = argc > 4 ? std::atoi(argv[4]) : 360;
    const uint32_t spp = argc > 5 ? std::atoi(argv[5]) : 16, bounces = argc > 6 ? std::atoi(argv[6]) : 4;
    flx_scene *scene = nullptr;
    flx_envmap *env = nullptr;
    try
    {
        if (flx_scene_load(argv[1], &scene) != 0)
            throw std::runtime_error(flx_io_last_error());
        if (argc > 7 && flx_envmap_load(argv[7], &env) != 0)
            throw std::runtime_error(flx_io_last_error());
        const uint32_t nTris = flx_scene_num_triangles(scene);
        std::vector<flx_Material> mats(flx_scene_materials(scene), flx_scene_materials(scene) + flx_scene_num_materials(scene));
        // textures: PNG and JPEG files are decoded by the library; a texture that cannot be read is left out and the materials
        // that use it fall back to their constants
        const std::string modelPath = argv[1];
        const std::string folder = modelPath.substr(0, modelPath.find_last_of('/') + 1);
        const uint32_t nTex = flx_scene_num_textures(scene);
        std::vector<uint8_t *> images(nTex, nullptr);
        std::vector<uint32_t> texW(nTex, 1), texH(nTex, 1);
        static uint8_t white[4] = {255, 255, 255, 255};
        uint32_t decoded = 0;
        for (uint32_t i = 0; i < nTex; i++)
        {
            if (flx_image_load((folder + flx_scene_texture_name(scene, i)).c_str(), &texW[i], &texH[i], &images[i]) == 0)
                decoded++;
            else
            {
                images[i] = nullptr;
                texW[i] = texH[i] = 1;
                for (auto &m : mats)
                {
                    if (m.map_Kd == (int)i) m.map_Kd = -1;
                    if (m.map_Ks == (int)i) m.map_Ks = -1;
                    if (m.map_N == (int)i) m.map_N = -1;
                }
            }
        }
        std::vector<const uint8_t *> imagePtrs(nTex);
        for (uint32_t i = 0; i < nTex; i++)
            imagePtrs[i] = images[i] ? images[i] : white;
        std::vector<flx_TexDescriptor> texDesc(nTex);
        size_t texBytes = 0;
        std::vector<uint8_t> texBlob;
        if (nTex)
        {
            if (flx_pack_textures(imagePtrs.data(), texW.data(), texH.data(), nTex, nullptr, nullptr, &texBytes) != 0)
                throw std::runtime_error(flx_io_last_error());
            texBlob.resize(texBytes);
            if (flx_pack_textures(imagePtrs.data(), texW.data(), texH.data(), nTex, texDesc.data(), texBlob.data(), &texBytes) != 0)
                throw std::runtime_error(flx_io_last_error());
        }
        for (uint8_t *im : images)
            flx_image_free(im);
        if (nTex)
            std::fprintf(stderr, "textures: %u of %u decoded%s\n", decoded, nTex, decoded < nTex ? "; the others fall back to material constants" : "");

        CLContext clctx(W * H);
        std::vector<flx_Node> nodes(2 * (size_t)nTris);
        std::vector<uint32_t> indices(nTris);
        uint32_t nNodes = 0;
        float buildMs = 0.0f;
        if (flx_build_bvh(clctx.handle(), flx_scene_triangles(scene), nTris, 8, FLX_BVH_PLOC_OPT, nodes.data(), (uint32_t)nodes.size(), &nNodes, indices.data(), &buildMs) != 0)
            throw std::runtime_error(flx_last_error(clctx.handle()));

        SceneArrays s;
        s.tris = flx_scene_triangles(scene); s.numTris = nTris;
        s.indices = indices.data(); s.numIndices = nTris;
        s.nodes = nodes.data(); s.numNodes = nNodes;
        s.materials = mats.data(); s.numMaterials = (uint32_t)mats.size();
        s.textures = nTex ? texDesc.data() : nullptr; s.numTextures = nTex;
        s.texData = nTex ? texBlob.data() : nullptr; s.texBytes = texBytes;
        clctx.uploadSceneData(s);
        if (env)
        {
            EnvMapArrays e;
            e.rgb = flx_envmap_rgb(env); e.width = flx_envmap_width(env); e.height = flx_envmap_height(env);
            e.probTable = flx_envmap_prob(env); e.aliasTable = flx_envmap_alias(env); e.pdfTable = flx_envmap_pdf(env);
            clctx.createEnvMap(e);
        }
        clctx.setupPixelStorage(W, H);

        const flx_float3 lo = nodes[0].bmin, hi = nodes[0].bmax;
        const flx_float3 c = f3(0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z));
        const flx_float3 d = f3(hi.x - lo.x, hi.y - lo.y, hi.z - lo.z);
        const float radius = 0.5f * std::sqrt(d.x * d.x + d.y * d.y + d.z * d.z); // Tracer::init, src/tracer.cpp:66-67
        RenderParams params;
        std::memset(&params, 0, sizeof params);
        params.camera.pos = f3(c.x + 0.35f * radius, c.y + 0.45f * radius, c.z + 1.9f * radius);
        params.camera.dir = norm(f3(c.x - params.camera.pos.x, c.y - params.camera.pos.y, c.z - params.camera.pos.z));
        params.camera.right = norm(cross(params.camera.dir, f3(0, 1, 0)));
        params.camera.up = cross(params.camera.right, params.camera.dir);
        params.camera.fov = 60.0f;
        params.camera.focalDist = 0.5f;
        params.areaLight.pos = f3(params.camera.pos.x - 0.01f * params.camera.dir.x, params.camera.pos.y - 0.01f * params.camera.dir.y, params.camera.pos.z - 0.01f * params.camera.dir.z);
        params.areaLight.N = params.camera.dir;
        params.areaLight.right = params.camera.right;
        params.areaLight.up = params.camera.up;
        params.areaLight.size = flx_float2{0.5f, 0.5f};
        const float E = 200.0f * radius * radius; // the reference's E = 200 suits its unit-sized scenes; scale with the scene
        params.areaLight.E = f3(E, E, E);
        params.ppParams.exposure = 1.0f;
        params.ppParams.tmOperator = 2;
        params.width = W;
        params.height = H;
        params.n_tris = nTris;
        params.useAreaLight = 1;
        params.useEnvMap = env ? 1 : 0;
        params.envMapStrength = argc > 8 ? (float)std::atof(argv[8]) : 1.0f;
        params.maxBounces = bounces;
        params.sampleImpl = params.sampleExpl = 1;
        params.worldRadius = radius;
        clctx.updateParams(params);

        clctx.enqueueResetKernel(params);
        if (flx_timer_begin(clctx.handle()) != 0)
            throw std::runtime_error(flx_last_error(clctx.handle()));
        clctx.renderSingleLoop(spp);
        float ms = 0.0f;
        if (flx_timer_end(clctx.handle(), &ms) != 0)
            throw std::runtime_error(flx_last_error(clctx.handle()));
        clctx.enqueuePostprocessKernel(params);
        clctx.finishQueue();
        clctx.saveImage(argv[2], params);
        flx_RenderStats64 st;
        if (flx_get_stats(clctx.handle(), &st) != 0)
            throw std::runtime_error(flx_last_error(clctx.handle()));
        std::printf("{\"triangles\": %u, \"bvh_nodes\": %u, \"bvh_build_ms\": %.3f, \"width\": %u, \"height\": %u, \"spp\": %u, \"render_ms\": %.3f, \"mrays_per_s\": %.1f, \"saved\": \"%s\"}\n",
                    nTris, nNodes, buildMs, W, H, spp, ms, (st.primaryRays + st.extensionRays + st.shadowRays) / (ms * 1e3), argv[2]);
    }
    catch (const std::runtime_error &e)
    {
        std::fprintf(stderr, "error: %s\n", e.what());
        flx_scene_free(scene);
        flx_envmap_free(env);
        return 1;
    }
    flx_scene_free(scene);
    flx_envmap_free(env);
    return 0;
}
