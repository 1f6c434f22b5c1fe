// flx_bsdf.cuh -- texture fetch, normal mapping, BSDF sample/eval/pdf and environment-map
// lookups used by the logic and material stages.
//
// What is computed follows the reference (file:line on each function); how it is organised is
// this repo's own: one `Surface` record loaded once per path, BSDF lobes selected by a
// compile-time type mask so each per-material-queue kernel only contains its own lobe.
#pragma once

#include "flx_device.cuh"

struct SceneView
{
    const flx_Triangle *tris;
    const flx_Material *materials;
    const flx_TexDescriptor *textures;
    const uint8_t *texData;
    const float4 *kdGamma; // per material: pow(Kd, 2.2) evaluated once at upload with the same flx_powf (bit-identical)
    // environment map (RGBA32F) + alias-method tables
    const float *envRGBA;
    int envW, envH;
    const float *probTable;
    const int32_t *aliasTable;
    const float *pdfTable;
};

struct Surface // the hit record as the BSDF code sees it (reference: Hit, src/geom.h:133-142)
{
    V3 P, N;
    float u, v;
    int tri;
};

struct Mat // reference: Material, src/geom.h:113-124
{
    V3 Kd, Ks, KdGamma;
    float Ns, Ni;
    int map_Kd, map_Ks, map_N, type;
};

FLX_DEV Mat load_material(const flx_Material *materials, int id, const float4 *kdGamma)
{
    // 80-byte records, 16-byte aligned: five 128-bit loads
    const float4 *p = reinterpret_cast<const float4 *>(materials + id);
    const float4 a = __ldg(p + 0), b = __ldg(p + 1), d = __ldg(p + 3), e = __ldg(p + 4);
    Mat m;
    m.Kd = v3(a.x, a.y, a.z);
    m.Ks = v3(b.x, b.y, b.z);
    const float4 g = __ldg(kdGamma + id);
    m.KdGamma = v3(g.x, g.y, g.z);
    m.Ns = d.x;
    m.Ni = d.y;
    m.map_Kd = __float_as_int(d.z);
    m.map_Ks = __float_as_int(d.w);
    m.map_N = __float_as_int(e.x);
    m.type = __float_as_int(e.y);
    return m;
}

// ---- textures: nearest texel of a packed RGBA8 blob, origin lower-left (reference: src/utils.cl:114-147)
FLX_DEV V3 read_texture(float u, float v, const flx_TexDescriptor *textures, int idx, const uint8_t *texData)
{
    const uint32_t off = __ldg(&textures[idx].offset), width = __ldg(&textures[idx].width), height = __ldg(&textures[idx].height);
    const float ux = u * (float)width, uy = v * (float)height;
    const float fx = floorf(ux), fy = floorf(uy);
    // the reference mixes int and uint here; C promotes to unsigned (utils.cl:118-119)
    const int tx = (int)((((uint32_t)(int)fx) % width + width) % width);
    const int ty = (int)((((uint32_t)(int)fy) % height + height) % height);
    int cx = (int)(((float)tx + ux) - fx);
    int cy = (int)(((float)ty + uy) - fy);
    cx = min(max(cx, 0), (int)(width - 1));
    cy = min(max(cy, 0), (int)(height - 1));
    const uint8_t *pix = texData + off + (uint32_t)cx * 4u + (uint32_t)cy * width * 4u;
    const uchar4 t = *reinterpret_cast<const uchar4 *>(pix); // offsets are multiples of 4 (clcontext.cpp:570-611)
    return v3((float)t.x / 255.0f, (float)t.y / 255.0f, (float)t.z / 255.0f);
}
FLX_DEV V3 mat_float3(V3 fallback, float u, float v, int idx, const SceneView &sc) // utils.cl:144-147
{
    return (idx != -1) ? read_texture(u, v, sc.textures, idx, sc.texData) : fallback;
}
FLX_DEV V3 mat_albedo(V3 fallback, float u, float v, int idx, const SceneView &sc) // utils.cl:136-141 (gamma 2.2)
{
    const V3 c = mat_float3(fallback, u, v, idx, sc);
    return v3(flx_powf(c.x, 2.2f), flx_powf(c.y, 2.2f), flx_powf(c.z, 2.2f));
}

// ---- tangent-space normal map (reference: src/utils.cl:149-182)
FLX_DEV V3 shading_normal(const Surface &s, const Mat &m, const SceneView &sc)
{
    if (m.map_N == -1)
        return s.N;
    V3 tn = mat_float3(v3(0.5f, 0.5f, 1.0f), s.u, s.v, m.map_N, sc);
    tn = 2.0f * tn - v3(1.0f);
    const float4 *t = reinterpret_cast<const float4 *>(sc.tris + s.tri);
    const float4 p0 = __ldg(t + 0), t0 = __ldg(t + 2), p1 = __ldg(t + 3), t1v = __ldg(t + 5), p2 = __ldg(t + 6), t2v = __ldg(t + 8);
    const V3 e1 = v3(p1.x, p1.y, p1.z) - v3(p0.x, p0.y, p0.z);
    const V3 e2 = v3(p2.x, p2.y, p2.z) - v3(p0.x, p0.y, p0.z);
    const float t1x = t1v.x - t0.x, t1y = t1v.y - t0.y, t2x = t2v.x - t0.x, t2y = t2v.y - t0.y;
    const float det = t1x * t2y - t1y * t2x;
    if (det == 0.0f)
        return s.N;
    const float invDet = 1.0f / det;
    const V3 T = norm3(invDet * (e1 * t2y - e2 * t1y));
    const V3 B = norm3(invDet * (e2 * t1x - e1 * t2x));
    V3 N;
    N.x = (T.x * tn.x + B.x * tn.y) + s.N.x * tn.z;
    N.y = (T.y * tn.x + B.y * tn.y) + s.N.y * tn.z;
    N.z = (T.z * tn.x + B.z * tn.y) + s.N.z * tn.z;
    return norm3(N);
}

// ---- geometry helpers (reference: src/utils.cl:32-58, 82-112)
FLX_DEV V3 reflect3(V3 d, V3 n) { return d - (2.0f * dot3(d, n)) * n; }
FLX_DEV V3 refract3(V3 wi, V3 n, float eta)
{
    const float iDotN = dot3(-wi, n);
    const float sin2I = fmaxf(0.0f, 1.0f - iDotN * iDotN);
    const float sin2T = eta * eta * sin2I;
    const float cosT = sqrtf(fmaxf(0.0f, 1.0f - sin2T));
    return wi * eta + n * (eta * iDotN - cosT);
}
FLX_DEV void ortho_basis(V3 N, V3 &a, V3 &b)
{
    if (N.x != N.y || N.x != N.z)
        a = v3(N.z - N.y, N.x - N.z, N.y - N.x);
    else
        a = v3(N.z - N.y, N.x + N.z, -N.y - N.x);
    a = norm3(a);
    b = cross3(N, a);
}
FLX_DEV V3 cos_sample_hemisphere(V3 n, uint32_t &seed, float &pdf)
{
    const float r1 = (2.0f * FLX_PI_F) * flx_rand(seed);
    const float r2 = flx_rand(seed);
    const float r2s = sqrtf(r2);
    V3 w = n;
    V3 u = (fabsf(w.x) > 0.1f) ? cross3(v3(0.0f, 1.0f, 0.0f), w) : cross3(v3(1.0f, 0.0f, 0.0f), w);
    u = norm3(u);
    V3 v = cross3(w, u);
    u = u * (flx_cosf(r1) * r2s);
    v = v * (flx_sinf(r1) * r2s);
    w = w * sqrtf(1.0f - r2);
    const V3 dir = (u + v) + w;
    pdf = dot3(n, dir) / FLX_PI_F;
    return dir;
}

// ---- Fresnel, unpolarised dielectric (reference: src/fresnel.cl:5-20)
FLX_DEV float fresnel_dielectric(float cosI, float etaI, float etaT)
{
    const float sinI = sqrtf(fmaxf(0.0f, 1.0f - cosI * cosI));
    const float sinT = etaI / etaT * sinI;
    const float cosT = sqrtf(fmaxf(0.0f, 1.0f - sinT * sinT));
    if (sinT >= 1.0f)
        return 1.0f;
    const float parl = ((etaT * cosI) - (etaI * cosT)) / ((etaT * cosI) + (etaI * cosT));
    const float perp = ((etaI * cosI) - (etaT * cosT)) / ((etaI * cosI) + (etaT * cosT));
    return 0.5f * (parl * parl + perp * perp);
}

// ---- Lambert (reference: src/diffuse.cl:9-26)
FLX_DEV V3 diffuse_value(const Surface &s, const Mat &m, const SceneView &sc)
{
    // untextured: the gamma-expanded albedo is a per-material constant (utils.cl:136-141 evaluates pow per call)
    const V3 kd = (m.map_Kd == -1) ? m.KdGamma : mat_albedo(m.Kd, s.u, s.v, m.map_Kd, sc);
    return kd * FLX_INV_PI_F;
}
FLX_DEV float diffuse_pdf(const Surface &s, V3 dirOut) { return dot3(s.N, dirOut) * FLX_INV_PI_F; }
FLX_DEV V3 diffuse_sample(const Surface &s, const Mat &m, const SceneView &sc, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    dirOut = cos_sample_hemisphere(s.N, seed, pdfW);
    return diffuse_value(s, m, sc);
}

// ---- GGX microfacet after Walter et al. 2007 with Smith G1 (reference: src/ggx.cl:12-292)
FLX_DEV float ggx_roughness(float Ns) { return sqrtf(2.0f / (2.0f + Ns)); }
FLX_DEV V3 ggx_sample_lobe(float alpha, V3 N, uint32_t &seed)
{
    V3 X, Y;
    ortho_basis(N, X, Y);
    const float r0 = flx_rand(seed);
    const float r1 = flx_rand(seed);
    const float theta = flx_atan2f(alpha * sqrtf(r0), sqrtf(1.0f - r0));
    const float phi = FLX_2PI_F * r1;
    const float sinT = flx_sinf(theta), cosT = flx_cosf(theta), sinP = flx_sinf(phi), cosP = flx_cosf(phi);
    return norm3(((X * sinT) * cosP + (Y * sinT) * sinP) + N * cosT);
}
FLX_DEV float ggx_G1(float alpha, V3 v, V3 n, V3 m)
{
    const float mDotV = dot3(m, v), nDotV = dot3(n, v);
    if (nDotV * mDotV <= 0.0f)
        return 0.0f;
    const float c2 = nDotV * nDotV;
    const float tanSq = (c2 > 0.0f) ? ((1.0f - c2) / c2) : 0.0f;
    return 2.0f / (1.0f + sqrtf(1.0f + alpha * alpha * tanSq));
}
FLX_DEV float ggx_G(float alpha, V3 wi, V3 wo, V3 n, V3 m) { return ggx_G1(alpha, wi, n, m) * ggx_G1(alpha, wo, n, m); }
FLX_DEV float ggx_D(float alpha, V3 n, V3 m)
{
    const float nDotM = dot3(n, m);
    if (nDotM <= 0.0f)
        return 0.0f;
    const float c2 = nDotM * nDotM;
    const float tanSq = (nDotM != 0.0f) ? ((1.0f - c2) / c2) : 0.0f;
    const float aSq = alpha * alpha;
    const float denom = FLX_PI_F * c2 * c2 * (aSq + tanSq) * (aSq + tanSq);
    return denom > 0.0f ? (aSq / denom) : 0.0f;
}
FLX_DEV float ggx_pdf_reflect(float alpha, V3 wo, V3 N, V3 H)
{
    const float nDotH = fabsf(dot3(N, H)), oDotH = fabsf(dot3(wo, H));
    const float jInv = 4.0f * oDotH;
    return jInv == 0.0f ? 0.0f : ggx_D(alpha, N, H) * nDotH / jInv;
}
FLX_DEV V3 ggx_reflect_value(const Surface &s, float Ni, float Ns, V3 KsFallback, int map_Ks, const SceneView &sc, V3 wi /*outwards*/, V3 wo, V3 H)
{
    const float alpha = ggx_roughness(Ns);
    const float iDotN = dot3(wi, s.N), oDotN = dot3(wo, s.N);
    const float F = (Ni > 1.0f) ? fresnel_dielectric(iDotN, 1.0f, Ni) : 1.0f;
    const V3 Ks = mat_float3(KsFallback, s.u, s.v, map_Ks, sc);
    const float D = ggx_D(alpha, s.N, H);
    const float G = ggx_G(alpha, wi, wo, s.N, H);
    const float den = 4.0f * iDotN * oDotN;
    return (den != 0.0f) ? (((Ks * F) * G) * D) / den : v3(0.0f);
}
FLX_DEV V3 ggx_reflect_sample(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    const V3 wi = dirIn * -1.0f;
    const float alpha = ggx_roughness(m.Ns);
    const V3 H = ggx_sample_lobe(alpha, s.N, seed);
    dirOut = reflect3(-wi, H);
    pdfW = ggx_pdf_reflect(alpha, dirOut, s.N, H);
    return ggx_reflect_value(s, m.Ni, m.Ns, m.Ks, m.map_Ks, sc, wi, dirOut, H);
}
FLX_DEV V3 ggx_reflect_eval(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    const V3 wi = dirIn * -1.0f;
    const V3 H = norm3(wi + dirOut);
    return ggx_reflect_value(s, m.Ni, m.Ns, m.Ks, m.map_Ks, sc, wi, dirOut, H);
}
FLX_DEV float ggx_reflect_pdf(const Surface &s, float Ns, V3 dirIn, V3 dirOut)
{
    const V3 wi = dirIn * -1.0f;
    const V3 H = norm3(wi + dirOut);
    return ggx_pdf_reflect(ggx_roughness(Ns), dirOut, s.N, H);
}
FLX_DEV float ggx_pdf_refract(float alpha, float etaI, float etaO, V3 wi, V3 wo, V3 N, V3 H)
{
    const float nDotH = fabsf(dot3(N, H)), iDotH = fabsf(dot3(wi, H)), oDotH = fabsf(dot3(wo, H));
    const float sj = etaI * iDotH + etaO * oDotH;
    return sj == 0.0f ? 0.0f : ggx_D(alpha, N, H) * nDotH * oDotH * etaO * etaO / (sj * sj);
}
// transmission term shared by sample and eval (reference: ggx.cl:189-218 and 246-271)
FLX_DEV V3 ggx_transmit_value(const Surface &s, const Mat &m, const SceneView &sc, float alpha, float etaI, float etaO, float F, float iDotN,
                              float oDotN, float iDotH, float oDotH, V3 wi, V3 wo, V3 Nn, V3 H)
{
    const float eta = etaI / etaO;
    V3 bsdf = v3(eta * eta);
    const V3 Ks = mat_float3(m.Ks, s.u, s.v, m.map_Ks, sc);
    bsdf = bsdf * Ks;
    const float denom = iDotN * oDotN * (etaI * iDotH + etaO * oDotH) * (etaI * iDotH + etaO * oDotH);
    if (denom == 0.0f)
        return v3(0.0f);
    const float focus = etaO * etaO * iDotH * oDotH / denom;
    const float D = ggx_D(alpha, Nn, H);
    const float G = ggx_G(alpha, wi, wo, Nn, H);
    return ((((1.0f - F) * bsdf) * D) * G) * focus;
}
FLX_DEV V3 ggx_refract_sample(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    const V3 wi = dirIn * -1.0f;
    const float raylen = len3(wi);
    const float alpha = ggx_roughness(m.Ns);
    float etaI = 1.0f, etaO = m.Ni;
    if (backface)
    {
        const float t = etaI;
        etaI = etaO;
        etaO = t;
    }
    const float iDotN = dot3(norm3(wi), s.N);
    V3 H = ggx_sample_lobe(alpha, s.N, seed);
    const float F = fresnel_dielectric(iDotN, etaI, etaO);
    if (flx_rand(seed) < F)
    {
        dirOut = raylen * reflect3(norm3(-wi), H);
        pdfW = ggx_pdf_reflect(alpha, dirOut, s.N, H);
        const float oDotN = dot3(dirOut, s.N);
        const float D = ggx_D(alpha, s.N, H);
        const float G = ggx_G(alpha, wi, dirOut, s.N, H);
        const float den = 4.0f * iDotN * oDotN;
        return v3((den != 0.0f) ? (F * G * D / den) : 0.0f);
    }
    const float eta = etaI / etaO;
    dirOut = raylen * refract3(norm3(-wi), s.N, eta);
    H = norm3(-(wi * etaI + dirOut * etaO));
    const V3 Nn = backface ? -s.N : s.N;
    pdfW = ggx_pdf_refract(alpha, etaI, etaO, wi, dirOut, Nn, H);
    const float iDotH = fabsf(dot3(norm3(wi), H)), oDotH = fabsf(dot3(dirOut, H));
    const float oDotN = dot3(dirOut, s.N);
    return ggx_transmit_value(s, m, sc, alpha, etaI, etaO, F, iDotN, oDotN, iDotH, oDotH, wi, dirOut, Nn, H);
}
FLX_DEV V3 ggx_refract_eval(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    const V3 wi = dirIn * -1.0f;
    const float alpha = ggx_roughness(m.Ns);
    float etaI = 1.0f, etaO = m.Ni;
    if (backface)
    {
        const float t = etaI;
        etaI = etaO;
        etaO = t;
    }
    const float iDotN = dot3(norm3(wi), s.N);
    const float oDotN = dot3(norm3(dirOut), s.N);
    const float F = fresnel_dielectric(iDotN, etaI, etaO);
    if (!backface)
    {
        const V3 H = norm3(wi + dirOut);
        const float D = ggx_D(alpha, s.N, H);
        const float G = ggx_G(alpha, wi, dirOut, s.N, H);
        const float den = 4.0f * iDotN * oDotN;
        return (den != 0.0f) ? v3(F * G * D / den) : v3(0.0f);
    }
    const V3 H = norm3(-(wi * etaI + dirOut * etaO));
    const float iDotH = fabsf(dot3(norm3(wi), H)), oDotH = fabsf(dot3(norm3(dirOut), H));
    return ggx_transmit_value(s, m, sc, alpha, etaI, etaO, F, iDotN, oDotN, iDotH, oDotH, wi, dirOut, -s.N, H);
}
FLX_DEV float ggx_refract_pdf(const Surface &s, const Mat &m, bool backface, V3 dirIn, V3 dirOut)
{
    const V3 wi = dirIn * -1.0f;
    const float alpha = ggx_roughness(m.Ns);
    if (!backface)
    {
        const V3 H = norm3(wi + dirOut);
        return ggx_pdf_reflect(alpha, dirOut, s.N, H);
    }
    const float etaI = m.Ni, etaO = 1.0f;
    const V3 H = norm3(-(wi * etaI + dirOut * etaO));
    return ggx_pdf_refract(alpha, etaI, etaO, wi, dirOut, -s.N, H);
}

// ---- glossy = Fresnel-blended Lambert base under a GGX coat (reference: src/glossy.cl:12-101)
FLX_DEV float ks_to_eta(V3 Ks)
{
    const float k = fminf(fmaxf((Ks.x + Ks.y + Ks.z) / 3.0f, 0.0f), 0.99f);
    return (sqrtf(k) + 1.0f) / (1.0f - sqrtf(k));
}
FLX_DEV V3 eta_to_ks(float eta)
{
    const float r = (eta > 0.0f) ? ((eta - 1.0f) / (eta + 1.0f)) : 0.0f;
    return v3(r * r);
}
FLX_DEV Mat glossy_effective(const Surface &s, const Mat &m, const SceneView &sc, bool useLength)
{
    Mat e = m;
    e.Ks = mat_float3(m.Ks, s.u, s.v, m.map_Ks, sc);
    e.Ni = (m.Ni > 0.0f) ? m.Ni : ks_to_eta(e.Ks);
    const bool zero = useLength ? (len3(e.Ks) == 0.0f) : is_zero3(e.Ks); // glossy.cl:36 vs :76
    if (zero)
        e.Ks = eta_to_ks(e.Ni);
    return e;
}
FLX_DEV V3 glossy_sample(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    const Mat e = glossy_effective(s, m, sc, false);
    const float cosTh = dot3(norm3(-dirIn), s.N);
    const float F = fresnel_dielectric(cosTh, 1.0f, e.Ni);
    float basePdf, coatPdf;
    V3 base, coat;
    if (flx_rand(seed) < F)
    {
        coat = ggx_reflect_sample(s, e, sc, dirIn, dirOut, coatPdf, seed);
        base = diffuse_value(s, e, sc);
        basePdf = diffuse_pdf(s, dirOut);
    }
    else
    {
        base = diffuse_sample(s, e, sc, dirOut, basePdf, seed);
        coat = ggx_reflect_eval(s, e, sc, dirIn, dirOut);
        coatPdf = ggx_reflect_pdf(s, e.Ns, dirIn, dirOut);
    }
    if (dot3(s.N, dirOut) < 1e-5f)
        return v3(0.0f); // pdfW deliberately left as the caller initialised it (the reference leaves it unset, glossy.cl:58-59)
    pdfW = (1.0f - F) * basePdf + F * coatPdf;
    return base * (1.0f - F) + coat;
}
FLX_DEV V3 glossy_eval(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    const Mat e = glossy_effective(s, m, sc, true);
    const V3 base = diffuse_value(s, e, sc);
    const V3 coat = ggx_reflect_eval(s, e, sc, dirIn, dirOut);
    const float cosTh = dot3(norm3(-dirIn), s.N);
    const float F = fresnel_dielectric(cosTh, 1.0f, e.Ni);
    return base * (1.0f - F) + coat;
}
FLX_DEV float glossy_pdf(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    const V3 Ks = mat_float3(m.Ks, s.u, s.v, m.map_Ks, sc);
    const float Ni = (m.Ni > 0.0f) ? m.Ni : ks_to_eta(Ks);
    const float basePdf = diffuse_pdf(s, dirOut);
    const float coatPdf = ggx_reflect_pdf(s, m.Ns, dirIn, dirOut);
    const float cosTh = dot3(norm3(-dirIn), s.N);
    const float F = fresnel_dielectric(cosTh, 1.0f, Ni);
    return (1.0f - F) * basePdf + F * coatPdf;
}

// ---- delta lobes (reference: src/ideal_reflection.cl:9-33, src/ideal_dielectric.cl:10-56)
FLX_DEV V3 mirror_sample(const Surface &s, const Mat &m, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW)
{
    const float len = len3(dirIn);
    dirOut = len * reflect3(norm3(dirIn), s.N);
    pdfW = 1.0f;
    const V3 ks = mat_float3(m.Ks, s.u, s.v, m.map_Ks, sc);
    const float cosO = dot3(norm3(dirOut), s.N);
    return (cosO != 0.0f) ? ks / cosO : v3(0.0f);
}
FLX_DEV V3 dielectric_sample(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    const float raylen = len3(dirIn);
    V3 bsdf = v3(1.0f);
    const float cosI = dot3(norm3(-dirIn), s.N);
    float n1 = 1.0f, n2 = m.Ni;
    if (backface)
    {
        const float t = n1;
        n1 = n2;
        n2 = t;
    }
    const float eta = n1 / n2;
    const float fr = fresnel_dielectric(cosI, n1, n2);
    if (flx_rand(seed) < fr)
        dirOut = raylen * reflect3(norm3(dirIn), s.N);
    else
    {
        dirOut = raylen * refract3(norm3(dirIn), s.N, eta);
        bsdf = bsdf * (eta * eta);
        bsdf = bsdf * mat_float3(m.Ks, s.u, s.v, m.map_Ks, sc);
    }
    pdfW = 1.0f;
    const float cosO = dot3(norm3(dirOut), s.N);
    return bsdf / cosO;
}

// ---- type dispatch with lobes compiled in by MASK (reference: src/bxdf_partial.cl:19-153)
template <int MASK> FLX_DEV V3 bxdf_eval(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    if ((MASK & FLX_BXDF_DIFFUSE) && m.type == FLX_BXDF_DIFFUSE) return diffuse_value(s, m, sc);
    if ((MASK & FLX_BXDF_GLOSSY) && m.type == FLX_BXDF_GLOSSY) return glossy_eval(s, m, sc, dirIn, dirOut);
    if ((MASK & FLX_BXDF_GGX_ROUGH_REFLECTION) && m.type == FLX_BXDF_GGX_ROUGH_REFLECTION) return ggx_reflect_eval(s, m, sc, dirIn, dirOut);
    if ((MASK & FLX_BXDF_GGX_ROUGH_DIELECTRIC) && m.type == FLX_BXDF_GGX_ROUGH_DIELECTRIC) return ggx_refract_eval(s, m, backface, sc, dirIn, dirOut);
    if ((MASK & FLX_BXDF_EMISSIVE) && m.type == FLX_BXDF_EMISSIVE) return v3(1.0f);
    return v3(0.0f); // delta lobes evaluate to zero
}
template <int MASK> FLX_DEV float bxdf_pdf(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 dirOut)
{
    if ((MASK & FLX_BXDF_DIFFUSE) && m.type == FLX_BXDF_DIFFUSE) return diffuse_pdf(s, dirOut);
    if ((MASK & FLX_BXDF_GLOSSY) && m.type == FLX_BXDF_GLOSSY) return glossy_pdf(s, m, sc, dirIn, dirOut);
    if ((MASK & FLX_BXDF_GGX_ROUGH_REFLECTION) && m.type == FLX_BXDF_GGX_ROUGH_REFLECTION) return ggx_reflect_pdf(s, m.Ns, dirIn, dirOut);
    if ((MASK & FLX_BXDF_GGX_ROUGH_DIELECTRIC) && m.type == FLX_BXDF_GGX_ROUGH_DIELECTRIC) return ggx_refract_pdf(s, m, backface, dirIn, dirOut);
    return 0.0f;
}
template <int MASK> FLX_DEV V3 bxdf_sample(const Surface &s, const Mat &m, bool backface, const SceneView &sc, V3 dirIn, V3 &dirOut, float &pdfW, uint32_t &seed)
{
    if ((MASK & FLX_BXDF_DIFFUSE) && m.type == FLX_BXDF_DIFFUSE) return diffuse_sample(s, m, sc, dirOut, pdfW, seed);
    if ((MASK & FLX_BXDF_GLOSSY) && m.type == FLX_BXDF_GLOSSY) return glossy_sample(s, m, sc, dirIn, dirOut, pdfW, seed);
    if ((MASK & FLX_BXDF_GGX_ROUGH_REFLECTION) && m.type == FLX_BXDF_GGX_ROUGH_REFLECTION) return ggx_reflect_sample(s, m, sc, dirIn, dirOut, pdfW, seed);
    if ((MASK & FLX_BXDF_IDEAL_REFLECTION) && m.type == FLX_BXDF_IDEAL_REFLECTION) return mirror_sample(s, m, sc, dirIn, dirOut, pdfW);
    if ((MASK & FLX_BXDF_GGX_ROUGH_DIELECTRIC) && m.type == FLX_BXDF_GGX_ROUGH_DIELECTRIC) return ggx_refract_sample(s, m, backface, sc, dirIn, dirOut, pdfW, seed);
    if ((MASK & FLX_BXDF_IDEAL_DIELECTRIC) && m.type == FLX_BXDF_IDEAL_DIELECTRIC) return dielectric_sample(s, m, backface, sc, dirIn, dirOut, pdfW, seed);
    if ((MASK & FLX_BXDF_EMISSIVE) && m.type == FLX_BXDF_EMISSIVE) return v3(1.0f);
    return v3(0.0f);
}

// ---- environment map: lat-long mapping, fp32 bilinear fetch, alias-method sampling
// (reference: src/env_map.cl:14-106)
FLX_DEV void direction_to_uv(V3 d, float &u, float &v)
{
    if (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f)
    {
        u = 0.0f;
        v = 0.0f;
        return;
    }
    const float uu = 1.0f + flx_atan2f(d.x, -d.z) / FLX_PI_F;
    const float r = fminf(fmaxf(d.y / len3(d), -1.0f), 1.0f);
    v = flx_acosf(r) / FLX_PI_F;
    u = uu * 0.5f;
}
FLX_DEV V3 uv_to_direction(float u, float v)
{
    const float phi = v * FLX_PI_F;
    const float theta = (u * 2.0f - 1.0f) * FLX_PI_F;
    const float sinPhi = flx_sinf(phi), cosPhi = flx_cosf(phi), sinTh = flx_sinf(theta), cosTh = flx_cosf(theta);
    return v3(sinPhi * sinTh, cosPhi, -sinPhi * cosTh);
}
FLX_DEV V3 env_eval_dir(const SceneView &sc, V3 d)
{
    float u, v, o[4];
    direction_to_uv(d, u, v);
    flx_bilinear_rgba(sc.envRGBA, sc.envW, sc.envH, u, v, o);
    return v3(o[0], o[1], o[2]);
}
FLX_DEV void env_sample_alias(const SceneView &sc, float rnd, V3 &L, float &pdfW)
{
    const int width = sc.envW, height = sc.envH;
    const float r = rnd * (float)width * (float)height;
    const int i = min((int)floorf(r), width * height - 1);
    const float mProb = __ldg(sc.probTable + i);
    const int uvInd = (r - (float)i < mProb) ? i : __ldg(sc.aliasTable + i);
    const float pdf_uv = __ldg(sc.pdfTable + uvInd);
    const int uInd = uvInd % width, vInd = uvInd / width;
    const float u = ((float)uInd + 0.5f) / (float)width;
    const float v = ((float)vInd + 0.5f) / (float)height;
    L = uv_to_direction(u, v);
    const float sinTh = flx_sinf(FLX_PI_F * v);
    const float directPdfUV = pdf_uv * 1.0f;
    pdfW = (sinTh != 0.0f) ? directPdfUV / (2.0f * FLX_PI_F * FLX_PI_F * sinTh) : 0.0f;
}
FLX_DEV float env_pdf(const SceneView &sc, V3 d)
{
    float u, v;
    direction_to_uv(d, u, v);
    const float sinTh = flx_sinf(v * FLX_PI_F);
    if (sinTh == 0.0f)
        return 0.0f;
    const int iu = min((int)floorf(u * (float)sc.envW), sc.envW - 1);
    const int iv = min((int)floorf(v * (float)sc.envH), sc.envH - 1);
    return __ldg(sc.pdfTable + iv * sc.envW + iu) / (FLX_2PI_F * FLX_PI_F * sinTh);
}
