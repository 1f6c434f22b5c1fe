// scene_tool.cpp -- TEST INFRASTRUCTURE (oracle).
// Runs the reference's own host-side data producers -- tinyobjloader (include/tiny_obj_loader.h),
// the SBVH builder (src/sbvh.cpp, src/bvh.cpp, src/bvhnode.cpp) and the environment-map
// importance tables (src/envmap.cpp, src/rgbe/rgbe.cpp) -- compiled unmodified against stub
// headers, and writes their outputs as flat binary blobs.  These blobs are INPUTS of the hot
// path; building/loading scenes is out of scope for the product.
//
// The OBJ->triangle/material conversion below follows src/scene.cpp:191-301 (matId = tinyobj
// id + 1, face normal when any vertex normal is missing, t = (u,v,0), "shader" MTL key ->
// BSDF type) and src/scene.cpp:13-26 (default material 0); the PLY path follows
// src/scene.cpp:422-553, 815-861.
#define TINYOBJLOADER_IMPLEMENTATION
#include "tiny_obj_loader.h"

#include "sbvh.hpp"
#include "envmap.hpp"
#include "progressview.hpp"
#include "geom.h"
#include "bxdf_types.h"

#include <array>
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

// no-op UI (src/progressview.hpp:13-24)
ProgressView::ProgressView(nanogui::Screen *) {}
void ProgressView::showError(const std::string &) {}
void ProgressView::showMessage(const std::string &, float) {}
void ProgressView::showMessage(const std::string &, const std::string &) {}
void ProgressView::showMessage(const std::string &, const std::string &, float) {}
void ProgressView::center() {}
void ProgressView::hide() {}

// bvh.hpp:15 declares "friend class CLContext": a class of that name may read the results.
class CLContext
{
  public:
    static const std::vector<Node> &nodes(const BVH &b) { return b.m_nodes; }
    static const std::vector<U32> &indices(const BVH &b) { return b.m_indices; }
};

static_assert(sizeof(RTTriangle) == 160, "RTTriangle layout");
static_assert(sizeof(Node) == 48, "Node layout");
static_assert(sizeof(Material) == 80, "Material layout");

static std::string unixify(std::string p)
{
    for (char &c : p)
        if (c == '\\')
            c = '/';
    return p;
}

static int shaderType(const std::string &type) // src/scene.cpp:171-189
{
    if (type == "diffuse") return BXDF_DIFFUSE;
    if (type == "glossy") return BXDF_GLOSSY;
    if (type == "rough_reflection") return BXDF_GGX_ROUGH_REFLECTION;
    if (type == "ideal_reflection") return BXDF_IDEAL_REFLECTION;
    if (type == "rough_dielectric") return BXDF_GGX_ROUGH_DIELECTRIC;
    if (type == "ideal_dielectric") return BXDF_IDEAL_DIELECTRIC;
    if (type == "emissive") return BXDF_EMISSIVE;
    return BXDF_DIFFUSE;
}

struct SceneData
{
    std::vector<RTTriangle> tris;
    std::vector<Material> mats;
    std::vector<std::string> texNames; // relative to the OBJ folder, first-use order (scene.cpp:304-321)
};

static int texIndex(SceneData &s, const std::string &name)
{
    if (name.empty())
        return -1;
    const std::string n = unixify(name);
    for (size_t i = 0; i < s.texNames.size(); i++)
        if (s.texNames[i] == n)
            return (int)i;
    s.texNames.push_back(n);
    return (int)s.texNames.size() - 1;
}

static void defaultMaterial(SceneData &s) // src/scene.cpp:13-26
{
    Material def;
    memset(&def, 0, sizeof def);
    def.Kd = fr::float3(0.64, 0.64, 0.64);
    def.Ni = 1.8f;
    def.Ns = 700.0f;
    def.map_Kd = def.map_Ks = def.map_N = -1;
    def.type = BXDF_DIFFUSE;
    s.mats.push_back(def);
}

static bool loadObj(const std::string &filePath, SceneData &s)
{
    std::vector<tinyobj::shape_t> shapes;
    std::vector<tinyobj::material_t> materials;
    tinyobj::attrib_t attrib;
    std::string err;
    size_t start = filePath.find_last_of("/");
    const std::string folder = filePath.substr(0, start + 1);
    if (!tinyobj::LoadObj(&attrib, &shapes, &materials, &err, filePath.c_str(), folder.c_str()))
    {
        std::cerr << "OBJ load failed: " << err << std::endl;
        return false;
    }
    const bool hasNormals = attrib.normals.size() > 0;
    const bool hasTexCoords = attrib.texcoords.size() > 0;
    for (auto &shape : shapes)
    {
        for (size_t f = 0; f < shape.mesh.indices.size() / 3; f++)
        {
            VertexPNT V[3];
            bool allNormals = true;
            for (size_t v = 0; v < 3; v++)
            {
                auto ind = shape.mesh.indices[3 * f + v];
                V[v].p = fr::float3(attrib.vertices[3 * ind.vertex_index + 0], attrib.vertices[3 * ind.vertex_index + 1],
                                    attrib.vertices[3 * ind.vertex_index + 2]);
                if (ind.normal_index < 0 || !hasNormals)
                {
                    allNormals = false;
                    V[v].n = fr::float3(0.0f);
                }
                else
                    V[v].n = fr::float3(attrib.normals[3 * ind.normal_index + 0], attrib.normals[3 * ind.normal_index + 1],
                                        attrib.normals[3 * ind.normal_index + 2]);
                if (ind.texcoord_index > -1 && hasTexCoords)
                    V[v].t = fr::float3(attrib.texcoords[2 * ind.texcoord_index + 0], attrib.texcoords[2 * ind.texcoord_index + 1], 0.0f);
                else
                    V[v].t = fr::float3(0.0f);
            }
            if (!allNormals)
                V[0].n = V[1].n = V[2].n = normalize(cross(V[1].p - V[0].p, V[2].p - V[0].p));
            RTTriangle tri(V[0], V[1], V[2]);
            tri.matId = shape.mesh.material_ids[f] + 1;
            s.tris.push_back(tri);
        }
    }
    for (auto &tm : materials)
    {
        Material m;
        memset(&m, 0, sizeof m);
        m.Kd = fr::float3(tm.diffuse[0], tm.diffuse[1], tm.diffuse[2]);
        m.Ks = fr::float3(tm.specular[0], tm.specular[1], tm.specular[2]);
        m.Ke = fr::float3(tm.emission[0], tm.emission[1], tm.emission[2]);
        m.Ns = tm.shininess;
        m.Ni = tm.ior;
        m.map_Kd = texIndex(s, tm.diffuse_texname);
        m.map_Ks = texIndex(s, tm.specular_texname);
        m.map_N = texIndex(s, tm.bump_texname);
        m.type = shaderType(tm.unknown_parameter["shader"]);
        s.mats.push_back(m);
    }
    return true;
}

static bool loadPly(const std::string &filename, SceneData &s) // ASCII PLY, src/scene.cpp:422-553
{
    struct Element
    {
        std::string name;
        int lines;
        std::vector<std::string> props;
    };
    std::vector<Element> elements;
    std::ifstream input(filename);
    if (!input)
        return false;
    std::string line, type = "none";
    int num = 0;
    std::vector<std::string> props;
    while (getline(input, line))
    {
        std::istringstream iss(line);
        std::string tok;
        iss >> tok;
        if (tok == "element")
        {
            elements.push_back(Element{type, num, props});
            props.clear();
            iss >> type >> num;
        }
        else if (tok == "property")
        {
            std::string t, n;
            iss >> t >> n;
            props.push_back(n);
        }
        else if (tok == "end_header")
        {
            elements.push_back(Element{type, num, props});
            break;
        }
    }
    std::vector<fr::float3> P, N;
    std::vector<std::array<unsigned, 3>> F;
    for (auto &e : elements)
    {
        for (int i = 0; i < e.lines; i++)
        {
            getline(input, line);
            std::istringstream iss(line);
            if (e.name == "vertex")
            {
                std::map<std::string, float> m;
                std::string b;
                for (auto &name : e.props)
                {
                    iss >> b;
                    m[name] = (float)atof(b.c_str());
                }
                P.push_back(fr::float3(m["x"], m["y"], m["z"]));
                if (m.find("nx") != m.end())
                    N.push_back(fr::float3(m["nx"], m["ny"], m["nz"]));
            }
            else if (e.name == "face")
            {
                int n;
                iss >> n;
                unsigned a, b, c, d;
                if (n == 3)
                {
                    iss >> a >> b >> c;
                    F.push_back({a, b, c});
                }
                else if (n == 4)
                {
                    iss >> a >> b >> c >> d;
                    F.push_back({a, b, c});
                    F.push_back({c, d, a});
                }
                else
                    return false;
            }
        }
    }
    for (auto &f : F) // src/scene.cpp:815-861, PLY branch
    {
        VertexPNT v0, v1, v2;
        v0.p = P[f[0]];
        v1.p = P[f[1]];
        v2.p = P[f[2]];
        if (N.empty())
            v0.n = v1.n = v2.n = normalize(cross(v1.p - v0.p, v2.p - v0.p));
        else
        {
            v0.n = N[f[0]];
            v1.n = N[f[1]];
            v2.n = N[f[2]];
        }
        s.tris.push_back(RTTriangle(v0, v1, v2));
    }
    return true;
}

template <class T> static void put(FILE *f, const T &v) { fwrite(&v, sizeof v, 1, f); }

int main(int argc, char **argv)
{
    if (argc < 4)
    {
        fprintf(stderr, "usage: scene_tool obj|ply|env <in> <out.bin>\n"
                        "       scene_tool cache-export obj|ply <model> <cache.bin>      (reference SBVH -> BVH::exportTo, src/bvh.cpp:174-192)\n"
                        "       scene_tool cache-import <cache.bin> <out.bin>            (BVH::importFrom, src/bvh.cpp:102-152 -> indices + 48-byte nodes)\n");
        return 2;
    }
    const std::string mode = argv[1], in = argv[2], out = argv[3];
    if (mode == "cache-export") // the reference's hierarchy cache file, written by the reference's own code
    {
        if (argc < 5)
            return 2;
        SceneData s;
        defaultMaterial(s);
        const bool ok = (in == "obj") ? loadObj(argv[3], s) : loadPly(argv[3], s);
        if (!ok || s.tris.empty())
            return 1;
        ProgressView pv(nullptr);
        SBVH bvh(&s.tris, SplitMode::SAH, &pv);
        bvh.exportTo(argv[4]);
        return 0;
    }
    if (mode == "cache-import") // ... and read back by the reference's own code
    {
        std::vector<RTTriangle> none;
        BVH bvh(&none, in);
        const auto &nodes = CLContext::nodes(bvh);
        const auto &idx = CLContext::indices(bvh);
        std::vector<Node> nodesOut(nodes);
        for (auto &n : nodesOut)
            memset((unsigned char *)&n + 41, 0, 7);
        FILE *f = fopen(out.c_str(), "wb");
        const uint32_t ni = idx.size(), nn = nodesOut.size();
        put(f, ni); put(f, nn);
        fwrite(idx.data(), sizeof(U32), ni, f);
        fwrite(nodesOut.data(), sizeof(Node), nn, f);
        fclose(f);
        return 0;
    }
    if (mode == "env")
    {
        EnvironmentMap env(in);
        if (!env.valid())
            return 1;
        FILE *f = fopen(out.c_str(), "wb");
        const uint32_t magic = 0x45584c46; // "FLXE"
        const uint32_t w = env.getWidth(), h = env.getHeight();
        put(f, magic); put(f, w); put(f, h);
        fwrite(env.getData(), sizeof(float), (size_t)w * h * 3, f);
        fwrite(env.getProbTable(), sizeof(float), (size_t)w * h, f);
        fwrite(env.getAliasTable(), sizeof(int), (size_t)w * h, f);
        fwrite(env.getPdfTable(), sizeof(float), (size_t)w * h, f);
        fclose(f);
        return 0;
    }
    SceneData s;
    defaultMaterial(s);
    bool ok = (mode == "obj") ? loadObj(in, s) : loadPly(in, s);
    if (!ok || s.tris.empty())
        return 1;
    // zero the struct padding so blobs are reproducible byte-for-byte
    for (auto &t : s.tris)
    {
        unsigned char *b = (unsigned char *)&t;
        memset(b + 148, 0, 12);
    }
    ProgressView pv(nullptr);
    SBVH bvh(&s.tris, SplitMode::SAH, &pv); // src/tracer.cpp:753-758
    const auto &nodes = CLContext::nodes(bvh);
    const auto &idx = CLContext::indices(bvh);
    std::vector<Node> nodesOut(nodes);
    for (auto &n : nodesOut) // zero padding bytes (41..47); nPrims at 40
        memset((unsigned char *)&n + 41, 0, 7);
    FILE *f = fopen(out.c_str(), "wb");
    const uint32_t magic = 0x53584c46; // "FLXS"
    const uint32_t nt = s.tris.size(), ni = idx.size(), nn = nodesOut.size(), nm = s.mats.size(), ntex = s.texNames.size();
    put(f, magic); put(f, nt); put(f, ni); put(f, nn); put(f, nm); put(f, ntex);
    fwrite(s.tris.data(), sizeof(RTTriangle), nt, f);
    fwrite(idx.data(), sizeof(U32), ni, f);
    fwrite(nodesOut.data(), sizeof(Node), nn, f);
    fwrite(s.mats.data(), sizeof(Material), nm, f);
    for (auto &n : s.texNames)
    {
        const uint32_t len = n.size();
        put(f, len);
        fwrite(n.data(), 1, len, f);
    }
    fclose(f);
    fprintf(stderr, "scene_tool: %u tris, %u indices, %u nodes, %u materials, %u textures\n", nt, ni, nn, nm, ntex);
    return 0;
}
