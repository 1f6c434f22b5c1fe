// flx_scene_io.cpp -- host-side scene input for the C ABI: Wavefront OBJ + MTL, ASCII PLY and Radiance
// RGBE environment maps into the arrays flx_upload_scene / flx_upload_envmap take, so a caller can go from files to a render
// without the reference's loaders.  Plain host C++, no device code.
//
// What is produced follows the reference (and is checked byte for byte against the reference's own loader code in tests/):
//   * OBJ/MTL: the reference loads through its vendored tinyobjloader 1.0.x (include/tiny_obj_loader.h) and converts in
//     Scene::loadObjWithMaterials (src/scene.cpp:191-301): triangle-fan triangulation in file order, matId = material + 1
//     (0 = the built-in default material, src/scene.cpp:13-26), face normal when any vertex lacks one, t = (u, v, 0), the
//     custom MTL key `shader` -> BSDF type (src/scene.cpp:171-189), textures numbered in first-use order
//     (src/scene.cpp:304-321).  Number parsing follows the loader's own decimal reader (tiny_obj_loader.h:463-586: digits
//     accumulated in a double, fraction digits scaled by a table / pow(10, -k), exponent through pow(5, e) and ldexp, then
//     rounded to float) -- NOT strtod, whose results differ in the last bit.
//   * PLY: Scene::loadPlyModel (src/scene.cpp:422-553): ASCII, element/property header, x y z [nx ny nz], faces of 3 or 4.
//   * RGBE: src/rgbe/rgbe.cpp:135-194 (header), 286-416 (run-length scanlines), 95-107 (rgbe -> float); importance tables:
//     EnvironmentMap::computeProbabilities, src/envmap.cpp:31-114 (luminance * sin(theta), pdf with mean 1, Vose's alias
//     method on two LIFO stacks).
// Image decoding (textures) is not done here: the reference uses DevIL, which is not available; callers decode the named
// files themselves and pass RGBA8 to flx_upload_scene.
#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "fluctus_b200.h"

bool flx_decode_jpeg(const std::string &path, uint32_t &w, uint32_t &h, std::vector<unsigned char> &rgba, std::string &error); // flx_jpeg.cpp

namespace
{
thread_local std::string g_io_error;

// ---- numbers the way the reference's OBJ loader reads them (tiny_obj_loader.h:463-586)
bool is_digit(char c) { return c >= '0' && c <= '9'; }

bool read_decimal(const char *s, const char *end, double *out)
{
    if (s >= end)
        return false;
    double mant = 0.0;
    int exp10 = 0;
    bool neg = false, expNeg = false;
    const char *p = s;
    if (*p == '+' || *p == '-')
    {
        neg = (*p == '-');
        p++;
    }
    else if (!is_digit(*p))
        return false;
    int nread = 0;
    while (p != end && is_digit(*p))
    {
        mant *= 10;
        mant += (int)(*p - '0');
        p++;
        nread++;
    }
    if (nread == 0)
        return false;
    bool haveExp = false;
    if (p != end)
    {
        if (*p == '.')
        {
            static const double lut[] = {1.0, 0.1, 0.01, 0.001, 0.0001, 0.00001, 0.000001, 0.0000001};
            p++;
            int k = 1;
            while (p != end && is_digit(*p))
            {
                mant += (int)(*p - '0') * (k < 8 ? lut[k] : std::pow(10.0, -k));
                k++;
                p++;
            }
            haveExp = (p != end) && (*p == 'e' || *p == 'E');
        }
        else
            haveExp = (*p == 'e' || *p == 'E');
    }
    if (haveExp)
    {
        p++;
        if (p != end && (*p == '+' || *p == '-'))
        {
            expNeg = (*p == '-');
            p++;
        }
        else if (!(p != end && is_digit(*p))) // the loader reads *p even at the end: a NUL or blank, i.e. "not a digit"
            return false;
        int digits = 0;
        while (p != end && is_digit(*p))
        {
            exp10 = exp10 * 10 + (int)(*p - '0');
            p++;
            digits++;
        }
        if (expNeg)
            exp10 = -exp10;
        if (digits == 0)
            return false;
    }
    *out = (neg ? -1 : 1) * (exp10 ? std::ldexp(mant * std::pow(5.0, exp10), exp10) : mant);
    return true;
}

float read_float(const char *&tok, double dflt = 0.0)
{
    tok += std::strspn(tok, " \t");
    const char *end = tok + std::strcspn(tok, " \t\r");
    double v = dflt;
    read_decimal(tok, end, &v);
    tok = end;
    return (float)v;
}

bool is_space(char c) { return c == ' ' || c == '\t'; }
bool is_eol(char c) { return c == '\r' || c == '\n' || c == '\0'; }

// getline that accepts \n, \r\n and \r (tiny_obj_loader.h safeGetline)
bool next_line(std::istream &in, std::string &line)
{
    line.clear();
    if (in.peek() == EOF)
        return false;
    std::streambuf *sb = in.rdbuf();
    for (;;)
    {
        const int c = sb->sbumpc();
        if (c == '\n')
            return true;
        if (c == '\r')
        {
            if (sb->sgetc() == '\n')
                sb->sbumpc();
            return true;
        }
        if (c == EOF)
        {
            if (line.empty())
                in.setstate(std::ios::eofbit);
            return true;
        }
        line += (char)c;
    }
}

struct Corner
{
    int v, vt, vn;
};
int fix_index(int idx, int n) { return idx > 0 ? idx - 1 : (idx == 0 ? 0 : n + idx); } // 1-based, negative = relative

Corner read_corner(const char *&tok, int nv, int nvn, int nvt) // i, i/j, i//k, i/j/k
{
    Corner c{-1, -1, -1};
    c.v = fix_index(std::atoi(tok), nv);
    tok += std::strcspn(tok, "/ \t\r");
    if (*tok != '/')
        return c;
    tok++;
    if (*tok == '/')
    {
        tok++;
        c.vn = fix_index(std::atoi(tok), nvn);
        tok += std::strcspn(tok, "/ \t\r");
        return c;
    }
    c.vt = fix_index(std::atoi(tok), nvt);
    tok += std::strcspn(tok, "/ \t\r");
    if (*tok != '/')
        return c;
    tok++;
    c.vn = fix_index(std::atoi(tok), nvn);
    tok += std::strcspn(tok, "/ \t\r");
    return c;
}

struct MtlEntry
{
    std::string name, mapKd, mapKs, mapBump;
    float Kd[3] = {0, 0, 0}, Ks[3] = {0, 0, 0}, Ke[3] = {0, 0, 0};
    float Ns = 1.0f, Ni = 1.0f; // tiny_obj_loader.h:860-861
    std::map<std::string, std::string> other;
};

std::string first_word(const char *tok) // sscanf("%s")
{
    tok += std::strspn(tok, " \t\r\n\v\f");
    return std::string(tok, std::strcspn(tok, " \t\r\n\v\f"));
}

// texture statement: options (-bm, -o, ...) are skipped with their arguments, the last bare word is the file name
// (tiny_obj_loader.h:746-840)
std::string texture_name(const char *tok)
{
    std::string name;
    while (!is_eol(*tok))
    {
        auto opt = [&](const char *o) {
            const size_t n = std::strlen(o);
            return std::strncmp(tok, o, n) == 0 && is_space(tok[n]);
        };
        auto skip_words = [&](int count) {
            for (int i = 0; i < count; i++)
            {
                tok += std::strspn(tok, " \t");
                tok += std::strcspn(tok, " \t\r");
            }
        };
        if (opt("-blendu") || opt("-blendv")) { tok += 8; skip_words(1); }
        else if (opt("-clamp") || opt("-boost")) { tok += 7; skip_words(1); }
        else if (opt("-bm")) { tok += 4; skip_words(1); }
        else if (opt("-o") || opt("-s") || opt("-t")) { tok += 3; skip_words(3); }
        else if (opt("-type")) { tok += 5; skip_words(1); }
        else if (opt("-imfchan")) { tok += 9; skip_words(1); }
        else if (opt("-mm")) { tok += 4; skip_words(2); }
        else
        {
            tok += std::strspn(tok, " \t");
            const size_t len = std::strcspn(tok, " \t\r");
            name.assign(tok, len);
            tok += len;
            tok += std::strspn(tok, " \t");
        }
    }
    return name;
}

void read_mtl(std::istream &in, std::vector<MtlEntry> &mats, std::map<std::string, int> &byName) // tiny_obj_loader.h:954-1318
{
    MtlEntry cur;
    std::string line;
    while (next_line(in, line))
    {
        if (!line.empty())
            line = line.substr(0, line.find_last_not_of(" \t") + 1);
        if (line.empty())
            continue;
        const char *tok = line.c_str();
        tok += std::strspn(tok, " \t");
        if (*tok == '\0' || *tok == '#')
            continue;
        auto key = [&](const char *k) {
            const size_t n = std::strlen(k);
            return std::strncmp(tok, k, n) == 0 && is_space(tok[n]);
        };
        if (key("newmtl"))
        {
            if (!cur.name.empty())
            {
                byName.insert(std::make_pair(cur.name, (int)mats.size())); // first definition of a name wins
                mats.push_back(cur);
            }
            cur = MtlEntry();
            cur.name = first_word(tok + 7);
            continue;
        }
        auto rgb = [&](float *dst) {
            tok += 2;
            dst[0] = read_float(tok);
            dst[1] = read_float(tok);
            dst[2] = read_float(tok);
        };
        if (key("Kd")) { rgb(cur.Kd); continue; }
        if (key("Ks")) { rgb(cur.Ks); continue; }
        if (key("Ke")) { rgb(cur.Ke); continue; }
        if (key("Ni")) { tok += 2; cur.Ni = read_float(tok); continue; }
        if (key("Ns")) { tok += 2; cur.Ns = read_float(tok); continue; }
        if (key("map_Kd")) { const std::string n = texture_name(tok + 7); if (!n.empty()) cur.mapKd = n; continue; }
        if (key("map_Ks")) { const std::string n = texture_name(tok + 7); if (!n.empty()) cur.mapKs = n; continue; }
        if (key("map_bump")) { const std::string n = texture_name(tok + 9); if (!n.empty()) cur.mapBump = n; continue; }
        if (key("bump")) { const std::string n = texture_name(tok + 5); if (!n.empty()) cur.mapBump = n; continue; }
        // statements the loader knows but the reference never reads
        static const char *ignored[] = {"Ka", "Kt", "Tf", "illum", "d", "Tr", "Pr", "Pm", "Ps", "Pc", "Pcr", "aniso", "anisor", "map_Ka", "map_Ns", "map_d", "disp",
                                        "map_Pr", "map_Pm", "map_Ps", "map_Ke", "norm"};
        bool known = false;
        for (const char *k : ignored)
            if (key(k))
                known = true;
        if (known)
            continue;
        // anything else is kept as key -> rest of the line; the reference reads "shader" from here
        const char *sp = std::strchr(tok, ' ');
        if (!sp)
            sp = std::strchr(tok, '\t');
        if (sp)
            cur.other.insert(std::make_pair(std::string(tok, sp - tok), std::string(sp + 1)));
    }
    byName.insert(std::make_pair(cur.name, (int)mats.size()));
    mats.push_back(cur);
}

int shader_type(const std::string &t) // src/scene.cpp:171-189
{
    if (t == "diffuse") return FLX_BXDF_DIFFUSE;
    if (t == "glossy") return FLX_BXDF_GLOSSY;
    if (t == "rough_reflection") return FLX_BXDF_GGX_ROUGH_REFLECTION;
    if (t == "ideal_reflection") return FLX_BXDF_IDEAL_REFLECTION;
    if (t == "rough_dielectric") return FLX_BXDF_GGX_ROUGH_DIELECTRIC;
    if (t == "ideal_dielectric") return FLX_BXDF_IDEAL_DIELECTRIC;
    if (t == "emissive") return FLX_BXDF_EMISSIVE;
    return FLX_BXDF_DIFFUSE;
}

flx_float3 f3(float x, float y, float z) { return flx_float3{x, y, z, 0.0f}; }
flx_float3 sub3(flx_float3 a, flx_float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
flx_float3 cross3(flx_float3 a, flx_float3 b) { return f3(a.y * b.z - b.y * a.z, b.x * a.z - a.x * b.z, a.x * b.y - a.y * b.x); } // include/math/float3.hpp:121
flx_float3 normalize3(flx_float3 v)                                                                                                // include/math/float3.hpp:41,45
{
    const float inv = 1.f / std::sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
    return f3(v.x * inv, v.y * inv, v.z * inv);
}
} // namespace

struct flx_scene
{
    std::vector<flx_Triangle> tris;
    std::vector<flx_Material> mats;
    std::vector<std::string> texNames;
};

struct flx_envmap
{
    int w = 0, h = 0;
    std::vector<float> rgb, prob, pdf;
    std::vector<int32_t> alias;
};

namespace
{
int tex_index(flx_scene &s, std::string name)
{
    if (name.empty())
        return -1;
    for (char &c : name)
        if (c == '\\')
            c = '/';
    for (size_t i = 0; i < s.texNames.size(); i++)
        if (s.texNames[i] == name)
            return (int)i;
    s.texNames.push_back(name);
    return (int)s.texNames.size() - 1;
}

void push_default_material(flx_scene &s) // src/scene.cpp:13-26
{
    flx_Material m;
    std::memset(&m, 0, sizeof m);
    m.Kd = f3(0.64f, 0.64f, 0.64f);
    m.Ni = 1.8f;
    m.Ns = 700.0f;
    m.map_Kd = m.map_Ks = m.map_N = -1;
    m.type = FLX_BXDF_DIFFUSE;
    s.mats.push_back(m);
}

flx_Triangle make_triangle(const flx_Vertex v[3], int matId)
{
    flx_Triangle t;
    std::memset(&t, 0, sizeof t);
    t.v0 = v[0];
    t.v1 = v[1];
    t.v2 = v[2];
    t.matId = matId;
    return t;
}

bool load_obj(const std::string &path, flx_scene &s)
{
    std::ifstream in(path.c_str());
    if (!in)
    {
        g_io_error = "cannot open " + path;
        return false;
    }
    size_t slash = path.find_last_of("\\");
    if (slash == std::string::npos)
        slash = path.find_last_of("/");
    const std::string folder = path.substr(0, slash + 1);

    std::vector<float> v, vn, vt;
    std::vector<MtlEntry> mtl;
    std::map<std::string, int> mtlByName;
    int material = -1;
    struct Face
    {
        Corner c[3];
        int material;
    };
    std::vector<Face> faces; // fan-triangulated, file order (shape boundaries do not reorder anything)
    std::string line;
    std::vector<Corner> poly;
    while (next_line(in, line))
    {
        if (line.empty())
            continue;
        const char *tok = line.c_str();
        tok += std::strspn(tok, " \t");
        if (*tok == '\0' || *tok == '#')
            continue;
        if (tok[0] == 'v' && is_space(tok[1]))
        {
            tok += 2;
            for (int k = 0; k < 3; k++)
                v.push_back(read_float(tok));
            continue;
        }
        if (tok[0] == 'v' && tok[1] == 'n' && is_space(tok[2]))
        {
            tok += 3;
            for (int k = 0; k < 3; k++)
                vn.push_back(read_float(tok));
            continue;
        }
        if (tok[0] == 'v' && tok[1] == 't' && is_space(tok[2]))
        {
            tok += 3;
            for (int k = 0; k < 2; k++)
                vt.push_back(read_float(tok));
            continue;
        }
        if (tok[0] == 'f' && is_space(tok[1]))
        {
            tok += 2;
            tok += std::strspn(tok, " \t");
            poly.clear();
            while (!is_eol(*tok))
            {
                poly.push_back(read_corner(tok, (int)(v.size() / 3), (int)(vn.size() / 3), (int)(vt.size() / 2)));
                tok += std::strspn(tok, " \t\r");
            }
            for (size_t k = 2; k < poly.size(); k++) // tiny_obj_loader.h:897-920
                faces.push_back(Face{{poly[0], poly[k - 1], poly[k]}, material});
            continue;
        }
        if (std::strncmp(tok, "usemtl", 6) == 0 && is_space(tok[6]))
        {
            const std::string name = first_word(tok + 7);
            const auto it = mtlByName.find(name);
            material = it != mtlByName.end() ? it->second : -1;
            continue;
        }
        if (std::strncmp(tok, "mtllib", 6) == 0 && is_space(tok[6]))
        {
            std::stringstream names(std::string(tok + 7));
            std::string one;
            while (std::getline(names, one, ' ')) // the first file that opens is used (tiny_obj_loader.h:1561-1576)
            {
                std::ifstream mf((folder + one).c_str());
                if (!mf)
                    continue;
                read_mtl(mf, mtl, mtlByName);
                break;
            }
            continue;
        }
        // g, o, s, t and unknown statements do not affect the triangle list
    }
    const bool hasNormals = !vn.empty(), hasTex = !vt.empty();
    const int nv = (int)(v.size() / 3), nvn = (int)(vn.size() / 3), nvt = (int)(vt.size() / 2);
    s.tris.reserve(faces.size());
    for (const Face &f : faces) // src/scene.cpp:229-283
    {
        flx_Vertex V[3];
        bool allNormals = true;
        for (int k = 0; k < 3; k++)
        {
            const Corner &c = f.c[k];
            if (c.v < 0 || c.v >= nv || c.vn >= nvn || c.vt >= nvt)
            {
                g_io_error = path + ": face index out of range";
                return false;
            }
            V[k].p = f3(v[3 * c.v], v[3 * c.v + 1], v[3 * c.v + 2]);
            if (c.vn < 0 || !hasNormals)
            {
                allNormals = false;
                V[k].n = f3(0, 0, 0);
            }
            else
                V[k].n = f3(vn[3 * c.vn], vn[3 * c.vn + 1], vn[3 * c.vn + 2]);
            V[k].t = (c.vt > -1 && hasTex) ? f3(vt[2 * c.vt], vt[2 * c.vt + 1], 0.0f) : f3(0, 0, 0);
        }
        if (!allNormals)
            V[0].n = V[1].n = V[2].n = normalize3(cross3(sub3(V[1].p, V[0].p), sub3(V[2].p, V[0].p)));
        s.tris.push_back(make_triangle(V, f.material + 1));
    }
    for (MtlEntry &m : mtl) // src/scene.cpp:286-301
    {
        flx_Material o;
        std::memset(&o, 0, sizeof o);
        o.Kd = f3(m.Kd[0], m.Kd[1], m.Kd[2]);
        o.Ks = f3(m.Ks[0], m.Ks[1], m.Ks[2]);
        o.Ke = f3(m.Ke[0], m.Ke[1], m.Ke[2]);
        o.Ns = m.Ns;
        o.Ni = m.Ni;
        o.map_Kd = tex_index(s, m.mapKd);
        o.map_Ks = tex_index(s, m.mapKs);
        o.map_N = tex_index(s, m.mapBump); // map_bump is treated as a normal map
        o.type = shader_type(m.other["shader"]);
        s.mats.push_back(o);
    }
    return true;
}

bool load_ply(const std::string &path, flx_scene &s) // src/scene.cpp:422-553, 815-861
{
    std::ifstream in(path.c_str());
    if (!in)
    {
        g_io_error = "cannot open " + path;
        return false;
    }
    struct Element
    {
        std::string name;
        int lines;
        std::vector<std::string> props;
    };
    // sizes in the header are claims: an element of n lines needs at least 2n bytes of file
    in.seekg(0, std::ios::end);
    const long long fileSize = (long long)in.tellg();
    in.seekg(0, std::ios::beg);
    std::vector<Element> elements;
    std::string line, type = "none";
    int count = 0;
    bool sawEnd = false;
    std::vector<std::string> props;
    while (std::getline(in, line))
    {
        std::istringstream iss(line);
        std::string tok;
        iss >> tok;
        if (tok == "format")
        {
            std::string fmt;
            iss >> fmt;
            if (fmt != "ascii") // the reference reads every PLY as text (src/scene.cpp:422-553); a binary body would parse as garbage
            {
                g_io_error = path + ": only ASCII PLY is supported (format is '" + fmt + "')";
                return false;
            }
        }
        else if (tok == "element")
        {
            elements.push_back(Element{type, count, props});
            props.clear();
            count = -1;
            iss >> type >> count;
            if (!iss || count < 0 || 2ll * count > fileSize)
            {
                g_io_error = path + ": element '" + type + "' claims a line count the file cannot hold";
                return false;
            }
        }
        else if (tok == "property")
        {
            std::string t, n;
            iss >> t >> n;
            props.push_back(n);
        }
        else if (tok == "end_header")
        {
            elements.push_back(Element{type, count, props});
            sawEnd = true;
            break;
        }
    }
    if (!sawEnd)
    {
        g_io_error = path + ": no end_header";
        return false;
    }
    std::vector<flx_float3> P, N;
    std::vector<unsigned> F;
    for (const Element &e : elements)
        for (int i = 0; i < e.lines; i++)
        {
            if (!std::getline(in, line))
            {
                g_io_error = path + ": body ends before element '" + e.name + "' is complete";
                return false;
            }
            std::istringstream iss(line);
            if (e.name == "vertex")
            {
                std::map<std::string, float> m;
                std::string word;
                for (const std::string &name : e.props)
                {
                    iss >> word;
                    m[name] = (float)std::atof(word.c_str());
                }
                P.push_back(f3(m["x"], m["y"], m["z"]));
                if (m.find("nx") != m.end())
                    N.push_back(f3(m["nx"], m["ny"], m["nz"]));
            }
            else if (e.name == "face")
            {
                int n = 0;
                iss >> n;
                unsigned a = 0, b = 0, c = 0, d = 0;
                if (n == 3)
                {
                    iss >> a >> b >> c;
                    F.insert(F.end(), {a, b, c});
                }
                else if (n == 4)
                {
                    iss >> a >> b >> c >> d;
                    F.insert(F.end(), {a, b, c, c, d, a});
                }
                else
                {
                    g_io_error = path + ": only faces of 3 or 4 vertices are supported";
                    return false;
                }
            }
        }
    if (!N.empty() && N.size() != P.size()) // e.g. two vertex elements of which one has normals: N[F[..]] would run off the end
    {
        g_io_error = path + ": some vertices have normals and some do not";
        return false;
    }
    for (size_t f = 0; f + 2 < F.size(); f += 3)
    {
        flx_Vertex V[3];
        std::memset(V, 0, sizeof V);
        for (int k = 0; k < 3; k++)
        {
            if (F[f + k] >= P.size())
            {
                g_io_error = path + ": face index out of range";
                return false;
            }
            V[k].p = P[F[f + k]];
        }
        if (N.empty())
            V[0].n = V[1].n = V[2].n = normalize3(cross3(sub3(V[1].p, V[0].p), sub3(V[2].p, V[0].p)));
        else
            for (int k = 0; k < 3; k++)
                V[k].n = N[F[f + k]];
        s.tris.push_back(make_triangle(V, 0));
    }
    return true;
}

// ---- Radiance RGBE (src/rgbe/rgbe.cpp)
bool read_rgbe(const std::string &path, flx_envmap &e)
{
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp)
    {
        g_io_error = "cannot open " + path;
        return false;
    }
    auto bail = [&](const char *why) {
        g_io_error = path + ": " + why;
        std::fclose(fp);
        return false;
    };
    char buf[128];
    if (!std::fgets(buf, sizeof buf, fp))
        return bail("empty file");
    // header lines up to the blank line, then the resolution line (rgbe.cpp:165-193)
    for (;;)
    {
        if (buf[0] == 0 || buf[0] == '\n')
            return bail("no FORMAT specifier found");
        if (!std::fgets(buf, sizeof buf, fp))
            return bail("truncated header");
        if (buf[0] == '\n')
            break;
    }
    if (!std::fgets(buf, sizeof buf, fp) || std::sscanf(buf, "-Y %d +X %d", &e.h, &e.w) < 2 || e.w <= 0 || e.h <= 0)
        return bail("missing image size specifier");
    const int w = e.w, h = e.h;
    {
        // The size line is a claim.  Flat data needs 4 bytes per pixel; a run-length scanline packs at most 127 bytes of a
        // channel into 2, so w * h pixels need at least w * h * 4 * 2 / 127 bytes.  (Also keeps w * h inside an int for the tables.)
        const long at = std::ftell(fp);
        std::fseek(fp, 0, SEEK_END);
        const long long rest = (long long)std::ftell(fp) - at;
        std::fseek(fp, at, SEEK_SET);
        const unsigned long long pixels = (unsigned long long)w * (unsigned long long)h;
        if (pixels > (1ull << 28) || pixels * 8ull / 127ull > (unsigned long long)std::max(rest, 0ll))
            return bail("image size specifier larger than the file can hold");
    }
    e.rgb.assign((size_t)w * h * 3, 0.0f);
    auto to_float = [](const unsigned char p[4], float *out) { // rgbe.cpp:95-107
        if (p[3])
        {
            const float f = (float)std::ldexp(1.0, (int)p[3] - (128 + 8));
            out[0] = p[0] * f;
            out[1] = p[1] * f;
            out[2] = p[2] * f;
        }
        else
            out[0] = out[1] = out[2] = 0.0f;
    };
    float *data = e.rgb.data();
    std::vector<unsigned char> scan((size_t)w * 4);
    auto read_flat = [&](size_t pixels) {
        unsigned char px[4];
        for (size_t i = 0; i < pixels; i++, data += 3)
        {
            if (std::fread(px, 4, 1, fp) < 1)
                return false;
            to_float(px, data);
        }
        return true;
    };
    if (w < 8 || w > 0x7fff)
    {
        if (!read_flat((size_t)w * h))
            return bail("truncated pixel data");
        std::fclose(fp);
        return true;
    }
    for (int y = 0; y < h; y++)
    {
        unsigned char head[4];
        if (std::fread(head, 4, 1, fp) < 1)
            return bail("truncated pixel data");
        if (head[0] != 2 || head[1] != 2 || (head[2] & 0x80))
        {
            // not run-length encoded: this pixel, then everything else flat (rgbe.cpp:317-324)
            to_float(head, data);
            data += 3;
            if (!read_flat((size_t)w * (h - y) - 1))
                return bail("truncated pixel data");
            break;
        }
        if ((((int)head[2]) << 8 | head[3]) != w)
            return bail("wrong scanline width");
        for (int ch = 0; ch < 4; ch++) // each channel separately (rgbe.cpp:331-360)
        {
            unsigned char *dst = scan.data() + (size_t)ch * w, *end = dst + w;
            while (dst < end)
            {
                unsigned char two[2];
                if (std::fread(two, 2, 1, fp) < 1)
                    return bail("truncated pixel data");
                if (two[0] > 128)
                {
                    int run = two[0] - 128;
                    if (run == 0 || run > end - dst)
                        return bail("bad scanline data");
                    while (run-- > 0)
                        *dst++ = two[1];
                }
                else
                {
                    int lit = two[0];
                    if (lit == 0 || lit > end - dst)
                        return bail("bad scanline data");
                    *dst++ = two[1];
                    if (--lit > 0)
                    {
                        if (std::fread(dst, (size_t)lit, 1, fp) < 1)
                            return bail("truncated pixel data");
                        dst += lit;
                    }
                }
            }
        }
        for (int x = 0; x < w; x++, data += 3)
        {
            const unsigned char px[4] = {scan[x], scan[(size_t)w + x], scan[2 * (size_t)w + x], scan[3 * (size_t)w + x]};
            to_float(px, data);
        }
    }
    std::fclose(fp);
    return true;
}

void importance_tables(flx_envmap &e) // src/envmap.cpp:31-114
{
    const int w = e.w, h = e.h, n = w * h;
    std::vector<float> scal((size_t)n);
    for (int v = 0; v < h; v++)
    {
        const float sinTh = std::sin(3.14159265358979323846f * float(v + 0.5f) / float(h));
        for (int u = 0; u < w; u++)
        {
            const float *p = &e.rgb[3 * ((size_t)v * w + u)];
            const float lum = 0.212671f * p[0] + 0.715160f * p[1] + 0.072169f * p[2];
            scal[(size_t)v * w + u] = lum * sinTh;
        }
    }
    e.pdf.assign((size_t)n, 0.0f);
    e.prob.assign((size_t)n, 0.0f);
    e.alias.assign((size_t)n, 0);
    float I = 0.0f;
    for (int i = 1; i < n + 1; i++)
        I += scal[i - 1] / (w * h);
    if (I == 0)
        for (int i = 0; i < n; i++)
            e.pdf[i] = 1.0f / float(n);
    else
        for (int i = 0; i < n; i++)
            e.pdf[i] = scal[i] / I;
    // Vose's alias method; both work lists are LIFO, which fixes which of the valid tables comes out
    std::vector<std::pair<float, int>> small, large;
    for (int i = 0; i < n; i++)
        (e.pdf[i] < 1.0f ? small : large).push_back(std::make_pair(e.pdf[i], i));
    while (!small.empty() && !large.empty())
    {
        const std::pair<float, int> l = small.back(), g = large.back();
        small.pop_back();
        large.pop_back();
        e.prob[l.second] = l.first;
        e.alias[l.second] = g.second;
        const float pg = (g.first + l.first) - 1.0f;
        (pg < 1.0f ? small : large).push_back(std::make_pair(pg, g.second));
    }
    for (auto &g : large)
        e.prob[g.second] = 1.0f;
    for (auto &l : small)
        e.prob[l.second] = 1.0f;
}

// ---- image output (CLContext::saveImage, src/clcontext.cpp:386-465, writes through DevIL; here two small writers)
uint32_t crc32_of(const unsigned char *p, size_t n, uint32_t crc)
{
    static uint32_t table[256];
    static bool ready = false;
    if (!ready)
    {
        for (uint32_t i = 0; i < 256; i++)
        {
            uint32_t c = i;
            for (int k = 0; k < 8; k++)
                c = (c & 1u) ? 0xedb88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; i++)
        crc = table[(crc ^ p[i]) & 0xffu] ^ (crc >> 8);
    return ~crc;
}

void put_be32(std::vector<unsigned char> &v, uint32_t x)
{
    for (int s = 24; s >= 0; s -= 8)
        v.push_back((unsigned char)(x >> s));
}

bool write_chunk(FILE *fp, const char type[4], const std::vector<unsigned char> &body)
{
    std::vector<unsigned char> head;
    put_be32(head, (uint32_t)body.size());
    std::vector<unsigned char> typed(type, type + 4);
    typed.insert(typed.end(), body.begin(), body.end());
    std::vector<unsigned char> tail;
    put_be32(tail, crc32_of(typed.data(), typed.size(), 0));
    return std::fwrite(head.data(), 1, 4, fp) == 4 && std::fwrite(typed.data(), 1, typed.size(), fp) == typed.size() && std::fwrite(tail.data(), 1, 4, fp) == 4;
}

// 8-bit RGB PNG; rows[0] is the BOTTOM row of the image (the renderer's y axis points up; the reference sets DevIL's origin
// to lower-left, src/main.cpp:69-71).  The zlib stream uses stored (uncompressed) deflate blocks: no dependency, valid PNG.
bool write_png(const std::string &path, const unsigned char *rgb, uint32_t w, uint32_t h)
{
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp)
    {
        g_io_error = "cannot create " + path;
        return false;
    }
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    bool ok = std::fwrite(sig, 1, 8, fp) == 8;
    std::vector<unsigned char> ihdr;
    put_be32(ihdr, w);
    put_be32(ihdr, h);
    const unsigned char fmt[5] = {8, 2, 0, 0, 0}; // 8 bits, truecolour, deflate, adaptive filtering (all rows filter 0), no interlace
    ihdr.insert(ihdr.end(), fmt, fmt + 5);
    ok = ok && write_chunk(fp, "IHDR", ihdr);
    std::vector<unsigned char> raw;
    raw.reserve((size_t)h * (3 * (size_t)w + 1));
    for (uint32_t y = 0; y < h; y++)
    {
        raw.push_back(0);
        const unsigned char *row = rgb + (size_t)(h - 1 - y) * w * 3;
        raw.insert(raw.end(), row, row + (size_t)w * 3);
    }
    std::vector<unsigned char> z;
    z.push_back(0x78);
    z.push_back(0x01);
    uint32_t a = 1, b = 0; // Adler-32
    for (size_t off = 0; off < raw.size() || off == 0; off += 65535)
    {
        const size_t n = std::min<size_t>(65535, raw.size() - off);
        z.push_back(off + n >= raw.size() ? 1 : 0);
        z.push_back((unsigned char)(n & 0xff));
        z.push_back((unsigned char)(n >> 8));
        z.push_back((unsigned char)(~n & 0xff));
        z.push_back((unsigned char)((~n >> 8) & 0xff));
        z.insert(z.end(), raw.begin() + off, raw.begin() + off + n);
        for (size_t i = 0; i < n; i++)
        {
            a = (a + raw[off + i]) % 65521u;
            b = (b + a) % 65521u;
        }
        if (raw.empty())
            break;
    }
    put_be32(z, (b << 16) | a);
    ok = ok && write_chunk(fp, "IDAT", z) && write_chunk(fp, "IEND", std::vector<unsigned char>());
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok)
        g_io_error = "write error on " + path;
    return ok;
}

// Radiance RGBE, flat (not run-length encoded) scanlines, top row first; rows[0] of the input is the bottom row
bool write_hdr(const std::string &path, const float *rgb, uint32_t w, uint32_t h)
{
    FILE *fp = std::fopen(path.c_str(), "wb");
    if (!fp)
    {
        g_io_error = "cannot create " + path;
        return false;
    }
    bool ok = std::fprintf(fp, "#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n-Y %u +X %u\n", h, w) > 0;
    std::vector<unsigned char> row((size_t)w * 4);
    for (uint32_t y = 0; y < h && ok; y++)
    {
        const float *src = rgb + (size_t)(h - 1 - y) * w * 3;
        for (uint32_t x = 0; x < w; x++)
        {
            const float r = src[3 * x], g = src[3 * x + 1], bl = src[3 * x + 2];
            float v = r > g ? r : g;
            if (bl > v)
                v = bl;
            unsigned char *o = &row[4 * (size_t)x];
            if (!(v >= 1e-32f)) // also catches NaN
                o[0] = o[1] = o[2] = o[3] = 0;
            else
            {
                int e;
                const float m = (float)(std::frexp(v, &e) * 256.0 / v);
                auto q = [&](float c) { const float t = c * m; return (unsigned char)(t <= 0.0f ? 0 : (t >= 255.0f ? 255 : (int)t)); };
                o[0] = q(r);
                o[1] = q(g);
                o[2] = q(bl);
                o[3] = (unsigned char)(e + 128);
            }
        }
        ok = std::fwrite(row.data(), 1, row.size(), fp) == row.size();
    }
    ok = (std::fclose(fp) == 0) && ok;
    if (!ok)
        g_io_error = "write error on " + path;
    return ok;
}

// ---- PNG decoding (textures).  The reference decodes images through DevIL (src/texture.cpp:16-40: ilLoadImage, then
// ilCopyPixels(... IL_RGBA, IL_UNSIGNED_BYTE ...) with the origin set to lower-left, src/main.cpp:69-71).  PNG is lossless, so any
// conforming decoder yields the same RGBA8 bytes; this one covers what DevIL's conversion to RGBA8 covers for PNG: 8- and 16-bit
// (high byte kept) grey, grey+alpha, RGB, RGBA and 1/2/4/8-bit palette or grey, tRNS transparency for palettes, all five
// scanline filters, non-interlaced.  JPEG is lossy and decoder-dependent ("texture decode parity unpinned"); it is not
// decoded here -- callers pass such textures decoded.
struct BitReader
{
    const unsigned char *p, *end;
    uint32_t acc = 0;
    int n = 0;
    bool bad = false;
    uint32_t bits(int k)
    {
        while (n < k)
        {
            if (p >= end)
            {
                bad = true;
                return 0;
            }
            acc |= (uint32_t)(*p++) << n;
            n += 8;
        }
        const uint32_t v = acc & ((k == 32) ? 0xffffffffu : ((1u << k) - 1u));
        acc >>= k;
        n -= k;
        return v;
    }
};

struct Huffman // canonical code, decoded bit by bit (count / symbol tables, RFC 1951 section 3.2.2)
{
    uint16_t count[16] = {0};
    uint16_t symbol[288] = {0};
    void build(const unsigned char *lengths, int nsym)
    {
        for (int i = 0; i < 16; i++)
            count[i] = 0;
        for (int i = 0; i < nsym; i++)
            count[lengths[i]]++;
        count[0] = 0;
        uint16_t offs[16];
        offs[1] = 0;
        for (int i = 1; i < 15; i++)
            offs[i + 1] = offs[i] + count[i];
        for (int i = 0; i < nsym; i++)
            if (lengths[i])
                symbol[offs[lengths[i]]++] = (uint16_t)i;
    }
    int decode(BitReader &br) const
    {
        int code = 0, first = 0, index = 0;
        for (int len = 1; len <= 15; len++)
        {
            code |= (int)br.bits(1);
            if (br.bad)
                return -1;
            const int c = count[len];
            if (code - c < first)
                return symbol[index + (code - first)];
            index += c;
            first += c;
            first <<= 1;
            code <<= 1;
        }
        return -1;
    }
};

// maxOut: the caller knows how much it expects (PNG: rows x (1 + row bytes)); a stream that inflates beyond it is rejected
// instead of being allowed to grow without bound (a 1 KB stream can expand ~1000-fold per level of nesting)
bool inflate_zlib(const unsigned char *src, size_t n, std::vector<unsigned char> &out, size_t maxOut)
{
    if (n < 6 || (src[0] & 0x0f) != 8 || ((src[0] << 8 | src[1]) % 31) != 0 || (src[1] & 0x20))
        return false;
    BitReader br{src + 2, src + n};
    static const uint16_t lenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
    static const uint16_t lenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
    static const uint16_t distBase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
    static const uint16_t distExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
    for (;;)
    {
        const uint32_t last = br.bits(1), type = br.bits(2);
        if (br.bad || type == 3)
            return false;
        if (type == 0)
        {
            br.acc = 0;
            br.n = 0; // to the byte boundary
            if (br.end - br.p < 4)
                return false;
            const uint32_t len = br.p[0] | (br.p[1] << 8), nlen = br.p[2] | (br.p[3] << 8);
            br.p += 4;
            if ((len ^ 0xffffu) != nlen || (size_t)(br.end - br.p) < len || out.size() + len > maxOut)
                return false;
            out.insert(out.end(), br.p, br.p + len);
            br.p += len;
        }
        else
        {
            Huffman lit, dist;
            unsigned char lengths[320];
            if (type == 1)
            {
                for (int i = 0; i < 288; i++)
                    lengths[i] = i < 144 ? 8 : (i < 256 ? 9 : (i < 280 ? 7 : 8));
                lit.build(lengths, 288);
                for (int i = 0; i < 30; i++)
                    lengths[i] = 5;
                dist.build(lengths, 30);
            }
            else
            {
                const int hlit = (int)br.bits(5) + 257, hdist = (int)br.bits(5) + 1, hclen = (int)br.bits(4) + 4;
                if (br.bad || hlit > 286 || hdist > 30)
                    return false;
                static const unsigned char order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
                unsigned char cl[19] = {0};
                for (int i = 0; i < hclen; i++)
                    cl[order[i]] = (unsigned char)br.bits(3);
                Huffman clh;
                clh.build(cl, 19);
                int i = 0;
                while (i < hlit + hdist)
                {
                    const int sym = clh.decode(br);
                    if (sym < 0)
                        return false;
                    if (sym < 16)
                        lengths[i++] = (unsigned char)sym;
                    else
                    {
                        int rep, val = 0;
                        if (sym == 16)
                        {
                            if (i == 0)
                                return false;
                            val = lengths[i - 1];
                            rep = 3 + (int)br.bits(2);
                        }
                        else if (sym == 17)
                            rep = 3 + (int)br.bits(3);
                        else
                            rep = 11 + (int)br.bits(7);
                        if (br.bad || i + rep > hlit + hdist)
                            return false;
                        while (rep--)
                            lengths[i++] = (unsigned char)val;
                    }
                }
                lit.build(lengths, hlit);
                dist.build(lengths + hlit, hdist);
            }
            for (;;)
            {
                const int sym = lit.decode(br);
                if (sym < 0)
                    return false;
                if (sym < 256)
                {
                    if (out.size() >= maxOut)
                        return false;
                    out.push_back((unsigned char)sym);
                }
                else if (sym == 256)
                    break;
                else
                {
                    if (sym > 285)
                        return false;
                    const int len = lenBase[sym - 257] + (int)br.bits(lenExtra[sym - 257]);
                    const int ds = dist.decode(br);
                    if (ds < 0 || ds > 29)
                        return false;
                    const size_t d = distBase[ds] + br.bits(distExtra[ds]);
                    if (br.bad || d > out.size() || out.size() + (size_t)len > maxOut)
                        return false;
                    const size_t from = out.size() - d;
                    for (int k = 0; k < len; k++)
                        out.push_back(out[from + k]);
                }
            }
        }
        if (last)
            return !br.bad;
    }
}

// -> RGBA8, row 0 = BOTTOM row (the reference's DevIL origin)
bool decode_png(const std::string &path, uint32_t &w, uint32_t &h, std::vector<unsigned char> &rgba)
{
    FILE *fp = std::fopen(path.c_str(), "rb");
    if (!fp)
    {
        g_io_error = "cannot open " + path;
        return false;
    }
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> file(size > 0 ? (size_t)size : 0);
    const bool readOk = file.empty() || std::fread(file.data(), 1, file.size(), fp) == file.size();
    std::fclose(fp);
    auto bad = [&](const char *why) {
        g_io_error = path + ": " + why;
        return false;
    };
    static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
    if (!readOk || file.size() < 8 + 25 || std::memcmp(file.data(), sig, 8) != 0)
        return bad("not a PNG file");
    auto be32 = [&](size_t at) { return ((uint32_t)file[at] << 24) | ((uint32_t)file[at + 1] << 16) | ((uint32_t)file[at + 2] << 8) | file[at + 3]; };
    int depth = 0, colour = 0, interlace = 0;
    std::vector<unsigned char> idat, palette, trns;
    bool haveHeader = false, done = false;
    for (size_t at = 8; at + 12 <= file.size() && !done;)
    {
        const uint32_t len = be32(at);
        if (len > file.size() - at - 12)
            return bad("truncated chunk");
        const char *type = reinterpret_cast<const char *>(&file[at + 4]);
        const unsigned char *body = &file[at + 8];
        if (crc32_of(&file[at + 4], (size_t)len + 4, 0) != be32(at + 8 + len))
            return bad("chunk checksum mismatch");
        if (!std::memcmp(type, "IHDR", 4))
        {
            if (len != 13)
                return bad("bad IHDR");
            w = be32(at + 8);
            h = be32(at + 12);
            depth = body[8];
            colour = body[9];
            interlace = body[12];
            if (body[10] != 0 || body[11] != 0)
                return bad("unknown compression / filter method");
            haveHeader = true;
        }
        else if (!std::memcmp(type, "PLTE", 4))
            palette.assign(body, body + len);
        else if (!std::memcmp(type, "tRNS", 4))
            trns.assign(body, body + len);
        else if (!std::memcmp(type, "IDAT", 4))
            idat.insert(idat.end(), body, body + len);
        else if (!std::memcmp(type, "IEND", 4))
            done = true;
        at += (size_t)len + 12;
    }
    if (!haveHeader || w == 0 || h == 0 || (uint64_t)w * h > (1ull << 28))
        return bad("missing or unreasonable header");
    if (interlace)
        return bad("interlaced PNG is not supported");
    int channels;
    switch (colour)
    {
    case 0: channels = 1; break;
    case 2: channels = 3; break;
    case 3: channels = 1; break;
    case 4: channels = 2; break;
    case 6: channels = 4; break;
    default: return bad("unknown colour type");
    }
    const bool depthOk = (colour == 0 && (depth == 1 || depth == 2 || depth == 4 || depth == 8 || depth == 16)) || (colour == 3 && (depth == 1 || depth == 2 || depth == 4 || depth == 8)) ||
                         ((colour == 2 || colour == 4 || colour == 6) && (depth == 8 || depth == 16));
    if (!depthOk)
        return bad("unsupported bit depth");
    if (colour == 3 && palette.size() < 3)
        return bad("palette image without a palette");
    std::vector<unsigned char> raw;
    const size_t rowBytes = ((size_t)w * channels * depth + 7) / 8, bpp = std::max<size_t>(1, (size_t)channels * depth / 8);
    raw.reserve((rowBytes + 1) * h);
    if (!inflate_zlib(idat.data(), idat.size(), raw, (rowBytes + 1) * h) || raw.size() < (rowBytes + 1) * h)
        return bad("corrupt image data");
    // undo the scanline filters in place (PNG specification, section 9)
    std::vector<unsigned char> zero(rowBytes, 0);
    for (uint32_t y = 0; y < h; y++)
    {
        unsigned char *cur = &raw[(rowBytes + 1) * y + 1];
        const unsigned char *up = y ? &raw[(rowBytes + 1) * (y - 1) + 1] : zero.data();
        const int filter = raw[(rowBytes + 1) * y];
        for (size_t x = 0; x < rowBytes; x++)
        {
            const int a = x >= bpp ? cur[x - bpp] : 0, b = up[x], c = x >= bpp ? up[x - bpp] : 0;
            int pred = 0;
            if (filter == 1) pred = a;
            else if (filter == 2) pred = b;
            else if (filter == 3) pred = (a + b) >> 1;
            else if (filter == 4)
            {
                const int pa = std::abs(b - c), pb = std::abs(a - c), pc = std::abs(a + b - 2 * c);
                pred = (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
            }
            else if (filter != 0)
                return bad("unknown scanline filter");
            cur[x] = (unsigned char)(cur[x] + pred);
        }
    }
    rgba.assign((size_t)w * h * 4, 255);
    for (uint32_t y = 0; y < h; y++)
    {
        const unsigned char *row = &raw[(rowBytes + 1) * y + 1];
        unsigned char *dst = &rgba[(size_t)(h - 1 - y) * w * 4]; // flip: row 0 of the result is the bottom row
        for (uint32_t x = 0; x < w; x++, dst += 4)
        {
            auto sample = [&](int ch) -> int { // channel value as 8 bits (16-bit: high byte; < 8-bit grey: scaled to 0..255)
                if (depth == 8)
                    return row[(size_t)x * channels + ch];
                if (depth == 16)
                    return row[((size_t)x * channels + ch) * 2];
                const size_t bit = (size_t)x * depth;
                const int v = (row[bit >> 3] >> (8 - depth - (bit & 7))) & ((1 << depth) - 1);
                return colour == 3 ? v : v * 255 / ((1 << depth) - 1);
            };
            if (colour == 3)
            {
                const size_t idx = (size_t)sample(0);
                if (idx * 3 + 2 >= palette.size())
                    return bad("palette index out of range");
                dst[0] = palette[idx * 3];
                dst[1] = palette[idx * 3 + 1];
                dst[2] = palette[idx * 3 + 2];
                dst[3] = idx < trns.size() ? trns[idx] : 255;
            }
            else if (colour == 0 || colour == 4)
            {
                dst[0] = dst[1] = dst[2] = (unsigned char)sample(0);
                dst[3] = colour == 4 ? (unsigned char)sample(1) : 255;
            }
            else
            {
                dst[0] = (unsigned char)sample(0);
                dst[1] = (unsigned char)sample(1);
                dst[2] = (unsigned char)sample(2);
                dst[3] = colour == 6 ? (unsigned char)sample(3) : 255;
            }
        }
    }
    return true;
}

bool ends_with(const std::string &s, const char *suffix)
{
    const size_t n = std::strlen(suffix);
    return s.size() >= n && s.compare(s.size() - n, n, suffix) == 0;
}
} // namespace

// No C++ exception may cross the C ABI (std::bad_alloc / std::length_error from a container sized by a hostile file would end
// in std::terminate inside the caller's process): every int-returning entry point is a function-try-block.
#define FLX_IO_CATCH                                                                                                   \
    catch (const std::exception &e)                                                                                    \
    {                                                                                                                  \
        g_io_error = std::string("out of memory or internal error: ") + e.what();                                      \
        return FLX_E_INVALID;                                                                                          \
    }                                                                                                                  \
    catch (...)                                                                                                        \
    {                                                                                                                  \
        g_io_error = "internal error";                                                                                 \
        return FLX_E_INVALID;                                                                                          \
    }

extern "C"
{
const char *flx_io_last_error(void) { return g_io_error.c_str(); }

int flx_scene_load(const char *path, flx_scene **out)
try
{
    if (!path || !out)
    {
        g_io_error = "flx_scene_load: null argument";
        return FLX_E_INVALID;
    }
    *out = nullptr;
    std::unique_ptr<flx_scene> owner(new flx_scene());
    flx_scene *s = owner.get();
    push_default_material(*s);
    const std::string p(path);
    bool ok = false;
    if (ends_with(p, "obj")) // Scene::loadModel dispatches on the file-name ending (src/scene.cpp:52-92)
        ok = load_obj(p, *s);
    else if (ends_with(p, "ply"))
        ok = load_ply(p, *s);
    else
        g_io_error = p + ": unsupported model format (obj and ply are)";
    if (ok && s->tris.empty())
    {
        g_io_error = p + ": no triangles";
        ok = false;
    }
    if (!ok)
        return FLX_E_INVALID;
    *out = owner.release();
    return 0;
}
FLX_IO_CATCH

void flx_scene_free(flx_scene *s) { delete s; }
uint32_t flx_scene_num_triangles(const flx_scene *s) { return s ? (uint32_t)s->tris.size() : 0; }
uint32_t flx_scene_num_materials(const flx_scene *s) { return s ? (uint32_t)s->mats.size() : 0; }
uint32_t flx_scene_num_textures(const flx_scene *s) { return s ? (uint32_t)s->texNames.size() : 0; }
const flx_Triangle *flx_scene_triangles(const flx_scene *s) { return s ? s->tris.data() : nullptr; }
const flx_Material *flx_scene_materials(const flx_scene *s) { return s ? s->mats.data() : nullptr; }
const char *flx_scene_texture_name(const flx_scene *s, uint32_t i) { return (s && i < s->texNames.size()) ? s->texNames[i].c_str() : nullptr; }

int flx_envmap_load(const char *path, flx_envmap **out)
try
{
    if (!path || !out)
    {
        g_io_error = "flx_envmap_load: null argument";
        return FLX_E_INVALID;
    }
    *out = nullptr;
    std::unique_ptr<flx_envmap> owner(new flx_envmap());
    flx_envmap *e = owner.get();
    if (!read_rgbe(path, *e))
        return FLX_E_INVALID;
    importance_tables(*e);
    *out = owner.release();
    return 0;
}
FLX_IO_CATCH

int flx_envmap_from_rgb(const float *rgb, int32_t w, int32_t h, flx_envmap **out)
try
{
    if (!rgb || !out || w <= 0 || h <= 0 || (unsigned long long)w * (unsigned long long)h > (1ull << 28))
    {
        g_io_error = "flx_envmap_from_rgb: bad arguments";
        return FLX_E_INVALID;
    }
    std::unique_ptr<flx_envmap> owner(new flx_envmap());
    flx_envmap *e = owner.get();
    e->w = w;
    e->h = h;
    e->rgb.assign(rgb, rgb + (size_t)w * h * 3);
    importance_tables(*e);
    *out = owner.release();
    return 0;
}
FLX_IO_CATCH

// CLContext::saveImage's two conversions (src/clcontext.cpp:407-451) on host buffers of n_pixels RGBA floats, row 0 = bottom row:
// *.hdr / *.HDR: the raw accumulator divided by its sample count, linear and unclamped, as Radiance RGBE;
// anything else: the post-processed preview (already tone-mapped and gamma-corrected) as 8-bit PNG, byte = (uchar)(255 * clamp01(c)).
int flx_write_image(const char *path, const float *rgba, uint32_t width, uint32_t height)
try
{
    if (!path || !rgba || width == 0 || height == 0)
    {
        g_io_error = "flx_write_image: bad arguments";
        return FLX_E_INVALID;
    }
    const std::string p(path);
    const size_t n = (size_t)width * height;
    if (ends_with(p, ".hdr") || ends_with(p, ".HDR"))
    {
        std::vector<float> rgb(n * 3);
        for (size_t i = 0; i < n; i++)
            for (int c = 0; c < 3; c++)
                rgb[3 * i + c] = rgba[4 * i + c] / rgba[4 * i + 3];
        return write_hdr(p, rgb.data(), width, height) ? 0 : FLX_E_INVALID;
    }
    std::vector<unsigned char> bytes(n * 3);
    for (size_t i = 0; i < n; i++)
        for (int c = 0; c < 3; c++)
            bytes[3 * i + c] = (unsigned char)(255 * std::max(0.0f, std::min(1.0f, rgba[4 * i + c])));
    return write_png(p, bytes.data(), width, height) ? 0 : FLX_E_INVALID;
}
FLX_IO_CATCH

// ---- the reference's hierarchy cache file (BVH::exportTo / importFrom, src/bvh.cpp:102-192; data/hierarchies/hierarchy_<hash>.bin,
// src/tracer.cpp:574-590): u32 nIndices, indices, u32 "node count", then per node 6 floats (box), u32 iStart/rightChild,
// i32 parent, u8 nPrims = 33 bytes.  The reference writes the INDEX count into the node-count field (src/bvh.cpp:185), so on
// reading the field is ignored and the node count is taken from the file length; on writing the true count goes in, which
// the reference's own importer reads correctly.
int flx_hierarchy_export(const char *path, const flx_Node *nodes, uint32_t n_nodes, const uint32_t *indices, uint32_t n_indices)
try
{
    if (!path || !nodes || !indices || n_nodes == 0 || n_indices == 0)
    {
        g_io_error = "flx_hierarchy_export: bad arguments";
        return FLX_E_INVALID;
    }
    FILE *fp = std::fopen(path, "wb");
    if (!fp)
    {
        g_io_error = std::string("cannot create ") + path;
        return FLX_E_INVALID;
    }
    std::vector<unsigned char> buf;
    buf.reserve(8 + (size_t)n_indices * 4 + (size_t)n_nodes * 33);
    auto put = [&](const void *p, size_t n) { buf.insert(buf.end(), (const unsigned char *)p, (const unsigned char *)p + n); };
    put(&n_indices, 4);
    put(indices, (size_t)n_indices * 4);
    put(&n_nodes, 4);
    for (uint32_t i = 0; i < n_nodes; i++)
    {
        const flx_Node &n = nodes[i];
        const float box[6] = {n.bmin.x, n.bmin.y, n.bmin.z, n.bmax.x, n.bmax.y, n.bmax.z};
        put(box, sizeof box);
        put(&n.iStartOrRightChild, 4);
        put(&n.parent, 4);
        put(&n.nPrims, 1);
    }
    const bool ok = std::fwrite(buf.data(), 1, buf.size(), fp) == buf.size();
    if (std::fclose(fp) != 0 || !ok)
    {
        g_io_error = std::string("write error on ") + path;
        return FLX_E_INVALID;
    }
    return 0;
}
FLX_IO_CATCH

// Two-call protocol: with nodes_out == NULL only the counts are returned; then call again with arrays of that size.
int flx_hierarchy_import(const char *path, flx_Node *nodes_out, uint32_t *n_nodes, uint32_t *indices_out, uint32_t *n_indices)
try
{
    if (!path || !n_nodes || !n_indices)
    {
        g_io_error = "flx_hierarchy_import: bad arguments";
        return FLX_E_INVALID;
    }
    FILE *fp = std::fopen(path, "rb");
    if (!fp)
    {
        g_io_error = std::string("cannot open ") + path;
        return FLX_E_INVALID;
    }
    std::fseek(fp, 0, SEEK_END);
    const long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    std::vector<unsigned char> buf(size > 0 ? (size_t)size : 0);
    const bool ok = buf.empty() || std::fread(buf.data(), 1, buf.size(), fp) == buf.size();
    std::fclose(fp);
    uint32_t ni = 0;
    if (ok && buf.size() >= 4)
        std::memcpy(&ni, buf.data(), 4);
    const size_t nodesAt = 4 + (size_t)ni * 4 + 4;
    if (!ok || buf.size() < 8 || ni == 0 || buf.size() < nodesAt + 33)
    {
        g_io_error = std::string(path) + ": not a hierarchy cache file";
        return FLX_E_INVALID;
    }
    const uint32_t nn = (uint32_t)((buf.size() - nodesAt) / 33);
    if (nodes_out && indices_out)
    {
        if (*n_nodes < nn || *n_indices < ni)
        {
            g_io_error = "flx_hierarchy_import: arrays too small";
            return FLX_E_INVALID;
        }
        std::memcpy(indices_out, buf.data() + 4, (size_t)ni * 4);
        for (uint32_t i = 0; i < nn; i++)
        {
            const unsigned char *p = buf.data() + nodesAt + (size_t)i * 33;
            flx_Node n;
            std::memset(&n, 0, sizeof n);
            float box[6];
            std::memcpy(box, p, 24);
            n.bmin = f3(box[0], box[1], box[2]);
            n.bmax = f3(box[3], box[4], box[5]);
            std::memcpy(&n.iStartOrRightChild, p + 24, 4);
            std::memcpy(&n.parent, p + 28, 4);
            n.nPrims = p[32];
            nodes_out[i] = n;
        }
    }
    *n_nodes = nn;
    *n_indices = ni;
    return 0;
}
FLX_IO_CATCH

// Decode one image file to RGBA8 with the reference's conventions (4 channels, row 0 = bottom row): PNG (decode_png above) and
// JPEG (flx_jpeg.cpp), the two formats the reference's scenes use.  The buffer belongs to the library until flx_image_free.
int flx_image_load(const char *path, uint32_t *width, uint32_t *height, uint8_t **rgba)
try
{
    if (!path || !width || !height || !rgba)
    {
        g_io_error = "flx_image_load: null argument";
        return FLX_E_INVALID;
    }
    *rgba = nullptr;
    std::string lower(path);
    for (char &c : lower)
        c = (char)std::tolower((unsigned char)c);
    std::vector<unsigned char> pixels;
    uint32_t w = 0, h = 0;
    if (ends_with(lower, ".png"))
    {
        if (!decode_png(path, w, h, pixels))
            return FLX_E_INVALID;
    }
    else if (ends_with(lower, ".jpg") || ends_with(lower, ".jpeg") || ends_with(lower, ".jpe"))
    {
        if (!flx_decode_jpeg(path, w, h, pixels, g_io_error))
            return FLX_E_INVALID;
    }
    else
    {
        g_io_error = std::string(path) + ": unsupported image format (PNG and JPEG are decoded)";
        return FLX_E_INVALID;
    }
    uint8_t *out = (uint8_t *)std::malloc(pixels.size());
    if (!out)
    {
        g_io_error = "flx_image_load: out of memory";
        return FLX_E_INVALID;
    }
    std::memcpy(out, pixels.data(), pixels.size());
    *rgba = out;
    *width = w;
    *height = h;
    return 0;
}
FLX_IO_CATCH

void flx_image_free(uint8_t *rgba) { std::free(rgba); }

// CLContext::packTextures (src/clcontext.cpp:570-611): descriptors {byte offset, width, height} + the RGBA8 images back to back.
// images[i] / widths[i] / heights[i]: the decoded textures in the scene's texture order.  Two-call protocol: blob_out == NULL
// returns the size needed.
int flx_pack_textures(const uint8_t *const *images, const uint32_t *widths, const uint32_t *heights, uint32_t n_tex, flx_TexDescriptor *desc_out, uint8_t *blob_out,
                      size_t *blob_bytes)
try
{
    if ((n_tex && (!images || !widths || !heights)) || !blob_bytes)
    {
        g_io_error = "flx_pack_textures: null argument";
        return FLX_E_INVALID;
    }
    size_t total = 0;
    for (uint32_t i = 0; i < n_tex; i++)
        total += (size_t)widths[i] * heights[i] * 4;
    if (total > 0xffffffffull)
    {
        g_io_error = "flx_pack_textures: more than 4 GiB of texture data (offsets are 32-bit, src/geom.h:126-131)";
        return FLX_E_INVALID;
    }
    if (blob_out && desc_out)
    {
        if (*blob_bytes < total)
        {
            g_io_error = "flx_pack_textures: blob too small";
            return FLX_E_INVALID;
        }
        size_t off = 0;
        for (uint32_t i = 0; i < n_tex; i++)
        {
            const size_t len = (size_t)widths[i] * heights[i] * 4;
            desc_out[i].offset = (uint32_t)off;
            desc_out[i].width = widths[i];
            desc_out[i].height = heights[i];
            std::memcpy(blob_out + off, images[i], len);
            off += len;
        }
    }
    *blob_bytes = total;
    return 0;
}
FLX_IO_CATCH

void flx_envmap_free(flx_envmap *e) { delete e; }
int32_t flx_envmap_width(const flx_envmap *e) { return e ? e->w : 0; }
int32_t flx_envmap_height(const flx_envmap *e) { return e ? e->h : 0; }
const float *flx_envmap_rgb(const flx_envmap *e) { return e ? e->rgb.data() : nullptr; }
const float *flx_envmap_prob(const flx_envmap *e) { return e ? e->prob.data() : nullptr; }
const int32_t *flx_envmap_alias(const flx_envmap *e) { return e ? e->alias.data() : nullptr; }
const float *flx_envmap_pdf(const flx_envmap *e) { return e ? e->pdf.data() : nullptr; }
}
