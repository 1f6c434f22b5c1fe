/*
 * wf_oracle.c -- TEST INFRASTRUCTURE: CPU restatement (oracle) of the reference's wavefront path.
 *
 * This file is the checker, not the product: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load the library built from it (oracle/liboracle.so).  Nothing in fluctus_b200/ does.
 *
 * It restates, in plain scalar C, what the reference's OpenCL kernels compute, one function per kernel, operating on
 * the reference's own buffer layouts (48-byte DFS nodes, u32 index list, 160-byte triangles, GPUTaskState SoA):
 *     port_reset     src/wf_reset.cl:5-66          port_logic_*   src/wf_logic.cl:14-314 (+ queue push 322-372)
 *     port_raygen    src/wf_raygen.cl:4-97         port_mat_*     src/wf_mat_*.cl + src/bxdf_partial.cl:19-153
 *     port_ext       src/wf_extrays.cl:5-36 + src/bvh.cl:234-310 + src/intersect.cl:41-155
 *     port_shadow    src/wf_shadowrays.cl:6-37 + src/bvh.cl:312-373
 *     port_mk_*      the microkernel integrator, src/mk_reset.cl, mk_raygen.cl, mk_next_vertex.cl, mk_sample_bsdf.cl, mk_splat.cl, mk_splat_preview.cl
 * Built-ins that OpenCL leaves implementation-defined are pinned exactly as in oracle/ref_shim/cl_shim.hpp
 * (include/flx_math.h for sin/cos/tan/atan2/acos/pow; IEEE 1/x, sqrt; dot = (xx'+yy')+zz'; normalize(0) = 0).
 *
 * Pinning: tests/test_oracle_cpu.py checks this restatement (a) against the golden vectors in tests/golden/ that were
 * produced by the reference's own kernel sources compiled for the host (oracle/_ref, generator tests/golden/make_golden.py)
 * and (b), where oracle/_ref is present, against those kernels directly on further configurations.  The reference itself
 * ships no tests or known-answer vectors for this path, so the pin is "reference source executed on the
 * host", not a vector published by the reference.
 *
 * The NDRange is executed serially in ascending work-item order (deterministic); with -DPORT_PARALLEL the loop is an
 * OpenMP parallel-for with real atomics (exported as port_par_*).
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "flx_math.h"
#include "ref_shim/ref_abi.h"

typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } f4; /* 16-byte float3 of the device layouts */

/* ---- device layouts (reference: src/geom.h:52-260) */
typedef struct { f4 bmin, bmax; int32_t parent; uint32_t link; uint8_t nPrims; uint8_t pad[7]; } Node;            /* geom.h:71-80 */
typedef struct { f4 p, n, t; } Vertex;
typedef struct { Vertex v0, v1, v2; int32_t matId; int32_t pad[3]; } Triangle;                                     /* geom.h:89-95 */
typedef struct { f4 Kd, Ks, Ke; float Ns, Ni; int32_t map_Kd, map_Ks, map_N, type; int32_t pad[2]; } Material;     /* geom.h:113-124 */
typedef struct { uint32_t offset, width, height; } TexDescriptor;
typedef struct { f4 right, up, N, pos, E; float sizex, sizey, pad[2]; } AreaLight;                                 /* geom.h:104-111 */
typedef struct { f4 pos, dir, up, right; float fov, apertureSize, focalDist, pad; } Camera;                        /* geom.h:146-155 */
typedef struct
{
    AreaLight areaLight; Camera camera; float exposure; uint32_t tmOperator;
    uint32_t width, height, n_tris, useEnvMap, useAreaLight; float envMapStrength;
    uint32_t maxBounces, sampleImpl, sampleExpl, useRoulette, wfSeparateQueues; float worldRadius; uint32_t pad[2];
} RenderParams;                                                                                                    /* geom.h:163-180 */
typedef struct { uint32_t raygen, extension, shadow, diffuse, glossy, ggxRefl, ggxRefr, delta; } QueueCounters;    /* geom.h:240-252 */
typedef char assert_sizes[(sizeof(Node) == 48 && sizeof(Triangle) == 160 && sizeof(Material) == 80 && sizeof(RenderParams) == 240) ? 1 : -1];

enum { BXDF_DIFFUSE = 2, BXDF_GLOSSY = 4, BXDF_GGX_REFL = 8, BXDF_IDEAL_REFL = 16, BXDF_GGX_REFR = 32, BXDF_IDEAL_DIEL = 64, BXDF_EMISSIVE = 128 }; /* bxdf_types.h */
#define IS_SINGULAR(t) (((t) & (BXDF_IDEAL_REFL | BXDF_IDEAL_DIEL)) != 0)

/* GPUTaskState SoA slots (geom.h:199-236 with 16-byte float3 members) */
enum { S_ORIG = 0, S_DIR = 4, S_SORIG = 8, S_SDIR = 12, S_T = 16, S_EI = 20, S_LBSDF = 24, S_LEMIT = 28, S_LT = 32, S_P = 36, S_N = 40, S_UV = 44,
       S_LPDFW = 47, S_LEN = 48, S_SEED = 49, S_LSPEC = 50, S_BLOCKED = 51, S_BACKFACE = 52, S_PIXEL = 53, S_FIRSTDIFF = 54, S_LPDFDIRECT = 55,
       S_LPDFIMPL = 56, S_LCOSTH = 57, S_LPICK = 58, S_SLEN = 59, S_HT = 60, S_HI = 61, S_HLIGHT = 62, S_HMAT = 63 };

typedef struct { uint32_t *base; uint32_t n; } Tasks;
static inline float rf(Tasks t, int s, uint32_t g) { float f; memcpy(&f, &t.base[(size_t)s * t.n + g], 4); return f; }
static inline uint32_t ru(Tasks t, int s, uint32_t g) { return t.base[(size_t)s * t.n + g]; }
static inline void wf(Tasks t, int s, uint32_t g, float f) { memcpy(&t.base[(size_t)s * t.n + g], &f, 4); }
static inline void wu(Tasks t, int s, uint32_t g, uint32_t u) { t.base[(size_t)s * t.n + g] = u; }
static inline v3 rv(Tasks t, int s, uint32_t g) { v3 r = {rf(t, s, g), rf(t, s + 1, g), rf(t, s + 2, g)}; return r; }
static inline void wv(Tasks t, int s, uint32_t g, v3 a) { wf(t, s, g, a.x); wf(t, s + 1, g, a.y); wf(t, s + 2, g, a.z); }

/* ---- vector arithmetic with the pinned operation order */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 V1(float s) { return V(s, s, s); }
static inline v3 F(f4 a) { return V(a.x, a.y, a.z); }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scl(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }   /* float3 * float */
static inline v3 lscl(float s, v3 a) { return V(s * a.x, s * a.y, s * a.z); }  /* float * float3 */
static inline v3 divs(v3 a, float s) { return V(a.x / s, a.y / s, a.z / s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline float dot(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length(v3 a) { return sqrtf(dot(a, a)); }
static inline v3 normalize(v3 a) { float l = length(a); if (l == 0.0f) return a; float i = 1.0f / l; return V(a.x * i, a.y * i, a.z * i); }
static inline int is_zero(v3 a) { return a.x == 0.0f && a.y == 0.0f && a.z == 0.0f; }
static inline v3 lerp3(float u, float v, v3 a, v3 b, v3 c) { return add(add(lscl(1.0f - u - v, a), lscl(u, b)), lscl(v, c)); } /* utils.cl:27-30 */
static inline float clampf(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* ---- RNG (random.cl:7-22) */
static inline uint32_t hash32(uint32_t s) { s = (s ^ 61u) ^ (s >> 16); s *= 9u; s = s ^ (s >> 4); s *= 0x27d4eb2du; s = s ^ (s >> 15); return s; }
static inline float rnd(uint32_t *seed) { *seed = hash32(*seed); return (float)(*seed) * (1.0f / 4294967296.0f); }

/* ---- atomics: serial by default */
#ifdef PORT_PARALLEL
#define NAME(n) port_par_##n
#define LOOP _Pragma("omp parallel for schedule(dynamic, 4096)") for (long long g_ = (long long)begin; g_ < (long long)end; ++g_)
static inline uint32_t atomic_inc(uint32_t *p) { return __atomic_fetch_add(p, 1u, __ATOMIC_RELAXED); }
static inline void add_float(float *p, float v)
{
    uint32_t old, neu;
    __atomic_load(( uint32_t *)p, &old, __ATOMIC_RELAXED);
    do { float f; memcpy(&f, &old, 4); f += v; memcpy(&neu, &f, 4); } while (!__atomic_compare_exchange_n((uint32_t *)p, &old, neu, 0, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
}
#else
#define NAME(n) port_##n
#define LOOP for (long long g_ = (long long)begin; g_ < (long long)end; ++g_)
static inline uint32_t atomic_inc(uint32_t *p) { uint32_t o = *p; *p = o + 1; return o; }
static inline void add_float(float *p, float v) { *p += v; }
#endif

static inline Tasks tasks_of(const RefBufs *b) { Tasks t = {(uint32_t *)b->tasks, b->numTasks}; return t; }

/* ================================================================ hit record helpers (utils.cl:202-224) */
typedef struct { v3 P, N; float u, v, t; int i, areaLightHit, matId; } Hit;
static inline Hit empty_hit(float tmax) { Hit h = {{0, 0, 0}, {0, 0, 0}, 0.0f, 0.0f, tmax, -1, 0, -1}; return h; } /* geom.h:144 */
static void write_hit(Tasks t, uint32_t g, const Hit *h)
{
    wv(t, S_P, g, h->P); wv(t, S_N, g, h->N); wf(t, S_UV, g, h->u); wf(t, S_UV + 1, g, h->v); wf(t, S_HT, g, h->t);
    wu(t, S_HI, g, (uint32_t)h->i); wu(t, S_HLIGHT, g, (uint32_t)h->areaLightHit); wu(t, S_HMAT, g, (uint32_t)h->matId);
}
static Hit read_hit(Tasks t, uint32_t g)
{
    Hit h; h.P = rv(t, S_P, g); h.N = rv(t, S_N, g); h.u = rf(t, S_UV, g); h.v = rf(t, S_UV + 1, g); h.t = rf(t, S_HT, g);
    h.i = (int)ru(t, S_HI, g); h.areaLightHit = (int)ru(t, S_HLIGHT, g); h.matId = (int)ru(t, S_HMAT, g); return h;
}
static void reset_path(Tasks t, uint32_t g, float worldRadius) /* the fields wf_reset.cl:31-56 and wf_raygen.cl:79-96 both reset */
{
    wv(t, S_EI, g, V1(0.0f)); wv(t, S_T, g, V1(1.0f)); wu(t, S_LEN, g, 0); wu(t, S_LSPEC, g, 1); wf(t, S_LPDFW, g, 1.0f);
    wf(t, S_LPDFDIRECT, g, 0.0f); wf(t, S_LPDFIMPL, g, 0.0f); wf(t, S_LCOSTH, g, 0.0f); wf(t, S_LPICK, g, 1.0f); wf(t, S_SLEN, g, 2.0f * worldRadius);
    wu(t, S_BACKFACE, g, 0); wu(t, S_BLOCKED, g, 1); wu(t, S_FIRSTDIFF, g, 0); wv(t, S_LEMIT, g, V1(0.0f)); wv(t, S_LBSDF, g, V1(0.0f));
    Hit h = empty_hit(FLT_MAX); write_hit(t, g, &h);
}

/* ================================================================ reset (wf_reset.cl:5-66) */
void NAME(reset)(const RefBufs *b, size_t begin, size_t end)
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); QueueCounters *ql = (QueueCounters *)b->queueLens;
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid < p->width * p->height)
        {
            for (int c = 0; c < 4; c++) { b->pixels[4 * (size_t)gid + c] = 0.0f; b->denoiserNormal[4 * (size_t)gid + c] = 0.0f; }
            b->denoiserAlbedo[4 * (size_t)gid + 0] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 1] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 2] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 3] = 0.0f;
        }
        if (gid >= b->numTasks) continue;
        reset_path(t, gid, p->worldRadius);
        wu(t, S_PIXEL, gid, 0); wu(t, S_SEED, gid, gid);
        b->raygenQueue[gid] = gid;
        if (gid == 0) ql->raygen = b->numTasks;
    }
}

/* ================================================================ raygen (wf_raygen.cl:4-97) */
void NAME(raygen)(const RefBufs *b, size_t begin, size_t end)
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); QueueCounters *ql = (QueueCounters *)b->queueLens;
    LOOP
    {
        const uint32_t gd = (uint32_t)g_;
        if (gd >= ql->raygen) continue;
        const uint32_t gid = b->raygenQueue[gd];
        uint32_t seed = ru(t, S_SEED, gid);
        const uint32_t numPixels = p->width * p->height;
        const uint32_t pixelIdx = (*b->currPixelIdx + gd) % numPixels;
        wu(t, S_PIXEL, gid, pixelIdx);
        float x = (float)(pixelIdx % p->width), y = (float)(pixelIdx / p->width);
        x += rnd(&seed); y += rnd(&seed);
        const float NDCx = x / p->width, NDCy = y / p->height;
        float SCRx = 2.0f * NDCx - 1.0f, SCRy = 2.0f * NDCy - 1.0f;
        SCRx *= (float)p->width / p->height;
        const float scale = flx_tanf(0.5f * p->camera.fov * 3.14159265358979323846f / 180); /* toRad, geom.h:22 */
        SCRx *= scale; SCRy *= scale;
        v3 rayOrig = F(p->camera.pos);
        const v3 target = add(add(add(rayOrig, scl(F(p->camera.right), SCRx)), scl(F(p->camera.up), SCRy)), F(p->camera.dir));
        v3 rayDir = normalize(sub(target, rayOrig));
        const v3 fp = add(F(p->camera.pos), scl(rayDir, p->camera.focalDist));
        const float sqrt_r = sqrtf(rnd(&seed)); const float th = FLX_2PI_F * rnd(&seed);                /* uniformSampleDisk, utils.cl:75-80 */
        const float rx = sqrt_r * flx_cosf(th), ry = sqrt_r * flx_sinf(th);
        rayOrig = add(rayOrig, lscl(p->worldRadius * p->camera.apertureSize, add(scl(F(p->camera.right), rx), scl(F(p->camera.up), ry))));
        rayDir = normalize(sub(fp, rayOrig));
        wv(t, S_ORIG, gid, rayOrig); wv(t, S_DIR, gid, rayDir);
        b->extensionQueue[atomic_inc(&ql->extension)] = gid;
        wu(t, S_SEED, gid, seed);
        reset_path(t, gid, p->worldRadius);
    }
}

/* ================================================================ traversal (bvh.cl:234-373, intersect.cl:41-155) */
typedef struct { unsigned long long V, B, T, U, rays; } WorkCount; /* SURVEY 8d terms; read with port_work_counts */
static WorkCount g_ext_work, g_shadow_work;
static int g_count_work = 0;
static uint32_t *g_visit_hist = 0; /* optional: per-node pop counts of the extension traversal (layout studies) */

static int intersect_aabb(v3 o, v3 d, const Node *box, float *tminRet, float tMaxPrev, WorkCount *wc)
{
    if (wc) wc->B++;
    const v3 dinv = V(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);                                             /* native_recip, intersect.cl:43 */
    const v3 tmp = mul(sub(F(box->bmin), o), dinv);
    v3 tmaxv = mul(sub(F(box->bmax), o), dinv);
    const v3 tminv = V(fminf(tmp.x, tmaxv.x), fminf(tmp.y, tmaxv.y), fminf(tmp.z, tmaxv.z));
    tmaxv = V(fmaxf(tmp.x, tmaxv.x), fmaxf(tmp.y, tmaxv.y), fmaxf(tmp.z, tmaxv.z));
    const float tmin = fmaxf(fmaxf(tminv.x, tminv.y), tminv.z), tmax = fminf(fminf(tmaxv.x, tmaxv.y), tmaxv.z);
    if (tmax < 0) return 0;
    if (tmin > tmax) return 0;
    *tminRet = tmin;
    return tmin < tMaxPrev;
}
static int intersect_tri_pts(v3 o, v3 d, v3 p0, v3 p1, v3 p2, float *tret, float *uret, float *vret)     /* intersect.cl:63-93 */
{
    const v3 s1 = sub(p1, p0), s2 = sub(p2, p0), pvec = cross(d, s2);
    const float det = dot(s1, pvec);
    if (fabsf(det) < 1e-12f) return 0;
    const float iDet = 1.0f / det;
    const v3 tvec = sub(o, p0);
    const float u = dot(tvec, pvec) * iDet;
    if (u < 0.0f || u > 1.0f) return 0;
    const v3 qvec = cross(tvec, s1);
    const float v = dot(d, qvec) * iDet;
    if (v < 0.0f || u + v > 1.0f) return 0;
    const float t = dot(s2, qvec) * iDet;
    if (t < 0.0f) return 0;
    *tret = t; *uret = u; *vret = v;
    return 1;
}
static int intersect_tri(v3 o, v3 d, const Triangle *tri, float *t, float *u, float *v, WorkCount *wc)
{
    if (wc) wc->T++;
    return intersect_tri_pts(o, d, F(tri->v0.p), F(tri->v1.p), F(tri->v2.p), t, u, v);
}
static void bvh_intersect(v3 o, v3 d, Hit *hit, const Triangle *tris, const Node *nodes, const uint32_t *indices, WorkCount *wc) /* bvh.cl:234-310 */
{
    uint32_t stack[64]; int sp = 0; stack[0] = 0;
    while (sp >= 0)
    {
        const uint32_t ni = stack[sp--];
        const Node *n = &nodes[ni];
        if (wc) wc->V++;
#ifndef PORT_PARALLEL
        if (g_visit_hist) g_visit_hist[ni]++;
#endif
        if (n->nPrims != 0)
        {
            float tmin = FLT_MAX, umin = 0.0f, vmin = 0.0f; int imin = -1;
            for (uint32_t i = n->link; i < n->link + n->nPrims; i++)
            {
                float t, u, v;
                if (intersect_tri(o, d, &tris[indices[i]], &t, &u, &v, wc) && t > 0.0f && t < tmin) { imin = (int)i; tmin = t; umin = u; vmin = v; }
            }
            if (imin != -1 && tmin < hit->t)
            {
                if (wc) wc->U++;
                const Triangle *T = &tris[indices[imin]];
                hit->i = (int)indices[imin]; hit->matId = T->matId; hit->t = tmin;
                hit->P = add(o, lscl(tmin, d));
                hit->N = normalize(lerp3(umin, vmin, F(T->v0.n), F(T->v1.n), F(T->v2.n)));
                const v3 uv = lerp3(umin, vmin, F(T->v0.t), F(T->v1.t), F(T->v2.t));
                hit->u = uv.x; hit->v = uv.y;
            }
        }
        else
        {
            float lnear = 0, rnear = 0;
            const int lh = intersect_aabb(o, d, &nodes[ni + 1], &lnear, hit->t, wc);
            const int rh = intersect_aabb(o, d, &nodes[n->link], &rnear, hit->t, wc);
            if (lh && rh)
            {
                uint32_t closer = ni + 1, farther = n->link;
                if (rnear < lnear) { const uint32_t tmp = closer; closer = farther; farther = tmp; }
                stack[++sp] = farther; stack[++sp] = closer;
            }
            else if (lh) stack[++sp] = ni + 1;
            else if (rh) stack[++sp] = n->link;
        }
    }
}
static int bvh_occluded(v3 o, v3 d, float maxDist, const Triangle *tris, const Node *nodes, const uint32_t *indices, WorkCount *wc) /* bvh.cl:312-373 */
{
    uint32_t stack[64]; int sp = 0; stack[0] = 0;
    while (sp >= 0)
    {
        const uint32_t ni = stack[sp--];
        const Node *n = &nodes[ni];
        if (wc) wc->V++;
        if (n->nPrims != 0)
        {
            for (uint32_t i = n->link; i < n->link + n->nPrims; i++)
            {
                float t, u, v;
                if (intersect_tri(o, d, &tris[indices[i]], &t, &u, &v, wc) && t > 0.0f && t < maxDist) return 1;
            }
        }
        else
        {
            float lnear = 0, rnear = 0;
            const int lh = intersect_aabb(o, d, &nodes[ni + 1], &lnear, maxDist, wc);
            const int rh = intersect_aabb(o, d, &nodes[n->link], &rnear, maxDist, wc);
            if (lh && rh)
            {
                uint32_t closer = ni + 1, farther = n->link;
                if (rnear < lnear) { const uint32_t tmp = closer; closer = farther; farther = tmp; }
                stack[++sp] = farther; stack[++sp] = closer;
            }
            else if (lh) stack[++sp] = ni + 1;
            else if (rh) stack[++sp] = n->link;
        }
    }
    return 0;
}
static int light_tri(v3 o, v3 d, v3 p0, v3 p1, v3 p2, float *tres)                                      /* intersectTriangleLocal, intersect.cl:96-121 */
{
    float t, u, v;
    if (!intersect_tri_pts(o, d, p0, p1, p2, &t, &u, &v)) return 0;
    if (t > *tres) return 0;
    *tres = t; return 1;
}
static void intersect_light(Hit *hit, v3 o, v3 d, const RenderParams *p)                                /* intersect.cl:124-155 */
{
    const AreaLight *L = &p->areaLight;
    if (dot(d, F(L->N)) > 0) return;
    const v3 pos = F(L->pos), right = F(L->right), up = F(L->up);
    const v3 tl = add(add(pos, lscl(L->sizex, right)), lscl(L->sizey, up)), tr = add(sub(pos, lscl(L->sizex, right)), lscl(L->sizey, up));
    const v3 bl = sub(add(pos, lscl(L->sizex, right)), lscl(L->sizey, up)), br = sub(sub(pos, lscl(L->sizex, right)), lscl(L->sizey, up));
    const int first = light_tri(o, d, tl, bl, br, &hit->t);
    const int second = light_tri(o, d, tl, br, tr, &hit->t);
    if (first || second) { hit->areaLightHit = 1; hit->P = add(o, lscl(hit->t, d)); hit->N = F(L->N); hit->i = 0; hit->matId = 0; }
}

void NAME(ext)(const RefBufs *b, size_t begin, size_t end)                                              /* wf_extrays.cl:5-36 */
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); const QueueCounters *ql = (const QueueCounters *)b->queueLens;
    LOOP
    {
        const uint32_t gd = (uint32_t)g_;
        if (gd >= ql->extension) continue;
        const uint32_t gid = b->extensionQueue[gd];
        const v3 o = rv(t, S_ORIG, gid), d = rv(t, S_DIR, gid);
        Hit hit = empty_hit(FLT_MAX);
        WorkCount wc = {0, 0, 0, 0, 1};
        bvh_intersect(o, d, &hit, (const Triangle *)b->tris, (const Node *)b->nodes, b->indices, g_count_work ? &wc : 0);
#ifndef PORT_PARALLEL
        if (g_count_work) { g_ext_work.V += wc.V; g_ext_work.B += wc.B; g_ext_work.T += wc.T; g_ext_work.U += wc.U; g_ext_work.rays += 1; }
#endif
        if (p->sampleImpl && p->useAreaLight) intersect_light(&hit, o, d, p);
        wu(t, S_LEN, gid, ru(t, S_LEN, gid) + 1);
        write_hit(t, gid, &hit);
    }
}

void NAME(shadow)(const RefBufs *b, size_t begin, size_t end)                                           /* wf_shadowrays.cl:6-37 */
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); const QueueCounters *ql = (const QueueCounters *)b->queueLens;
    LOOP
    {
        const uint32_t gd = (uint32_t)g_;
        if (gd >= ql->shadow) continue;
        const uint32_t gid = b->shadowQueue[gd];
        const v3 o = rv(t, S_SORIG, gid), d = rv(t, S_SDIR, gid);
        const float lenL = rf(t, S_SLEN, gid);
        Hit hitL = empty_hit(lenL);
        if (p->useAreaLight) intersect_light(&hitL, o, d, p);
        WorkCount wc = {0, 0, 0, 0, 1};
        const int occluded = (hitL.i > -1) || bvh_occluded(o, d, lenL, (const Triangle *)b->tris, (const Node *)b->nodes, b->indices, g_count_work ? &wc : 0);
#ifndef PORT_PARALLEL
        if (g_count_work) { g_shadow_work.V += wc.V; g_shadow_work.B += wc.B; g_shadow_work.T += wc.T; g_shadow_work.rays += 1; }
#endif
        wu(t, S_BLOCKED, gid, (uint32_t)occluded);
    }
}

#ifndef PORT_PARALLEL
void port_visit_hist(uint32_t *hist) { g_visit_hist = hist; }
void port_count_work(int enable) { g_count_work = enable; memset(&g_ext_work, 0, sizeof g_ext_work); memset(&g_shadow_work, 0, sizeof g_shadow_work); }
void port_work_counts(unsigned long long ext[5], unsigned long long shadow[5])
{
    ext[0] = g_ext_work.V; ext[1] = g_ext_work.B; ext[2] = g_ext_work.T; ext[3] = g_ext_work.U; ext[4] = g_ext_work.rays;
    shadow[0] = g_shadow_work.V; shadow[1] = g_shadow_work.B; shadow[2] = g_shadow_work.T; shadow[3] = 0; shadow[4] = g_shadow_work.rays;
}
#endif

/* ================================================================ textures, normal map (utils.cl:114-182) */
typedef struct { const Triangle *tris; const Material *materials; const TexDescriptor *textures; const uint8_t *texData;
                 const float *envRGBA; int envW, envH; const float *prob; const int32_t *alias; const float *pdf; } Scene;
static Scene scene_of(const RefBufs *b)
{
    Scene s = {(const Triangle *)b->tris, (const Material *)b->materials, (const TexDescriptor *)b->textures, b->texData, b->envRGBA, b->envW, b->envH,
               b->probTable, b->aliasTable, b->pdfTable};
    return s;
}
static v3 read_texture(float u, float v, TexDescriptor tex, const uint8_t *data)
{
    float ux = u * tex.width, uy = v * tex.height;
    const int tx = ((int)(floorf(ux)) % tex.width + tex.width) % tex.width;   /* int % uint -> unsigned arithmetic, as in the reference */
    const int ty = ((int)(floorf(uy)) % tex.height + tex.height) % tex.height;
    int cx = (int)(tx + ux - floorf(ux)), cy = (int)(ty + uy - floorf(uy));
    const int mx = (int)(tex.width - 1), my = (int)(tex.height - 1);
    cx = cx < 0 ? 0 : (cx > mx ? mx : cx); cy = cy < 0 ? 0 : (cy > my ? my : cy);
    const uint8_t *pix = data + tex.offset + cx * 4 + cy * tex.width * 4;
    v3 c = V((float)pix[0], (float)pix[1], (float)pix[2]);
    return divs(c, 255.0f);
}
static v3 mat_float3(v3 fallback, float u, float v, int idx, const Scene *sc) { return (idx != -1) ? read_texture(u, v, sc->textures[idx], sc->texData) : fallback; }
static v3 mat_albedo(v3 fallback, float u, float v, int idx, const Scene *sc)
{
    const v3 c = mat_float3(fallback, u, v, idx, sc);
    return V(flx_powf(c.x, 2.2f), flx_powf(c.y, 2.2f), flx_powf(c.z, 2.2f));
}
static v3 tangent_space_normal(const Hit *hit, const Material *mat, const Scene *sc)
{
    if (mat->map_N == -1) return hit->N;
    v3 tn = mat_float3(V(0.5f, 0.5f, 1.0f), hit->u, hit->v, mat->map_N, sc);
    tn = sub(lscl(2.0f, tn), V1(1.0f));
    const Triangle *t = &sc->tris[hit->i];
    const v3 e1 = sub(F(t->v1.p), F(t->v0.p)), e2 = sub(F(t->v2.p), F(t->v0.p));
    const v3 t1 = sub(F(t->v1.t), F(t->v0.t)), t2 = sub(F(t->v2.t), F(t->v0.t));
    const float det = t1.x * t2.y - t1.y * t2.x;
    if (det == 0.0) return hit->N;
    const float invDet = 1.0f / det;
    const v3 T = normalize(lscl(invDet, sub(scl(e1, t2.y), scl(e2, t1.y))));
    const v3 B = normalize(lscl(invDet, sub(scl(e2, t1.x), scl(e1, t2.x))));
    v3 N;
    N.x = T.x * tn.x + B.x * tn.y + hit->N.x * tn.z; N.y = T.y * tn.x + B.y * tn.y + hit->N.y * tn.z; N.z = T.z * tn.x + B.z * tn.y + hit->N.z * tn.z;
    return normalize(N);
}

/* ================================================================ environment map (env_map.cl:14-106) */
static void direction_to_uv(v3 d, float *u, float *v)
{
    if (d.x == 0.0f && d.y == 0.0f && d.z == 0.0f) { *u = 0.0f; *v = 0.0f; return; }
    const float uu = 1.0f + flx_atan2f(d.x, -d.z) / FLX_PI_F;
    const float r = clampf(d.y / length(d), -1.0f, 1.0f);
    *v = flx_acosf(r) / FLX_PI_F; *u = uu * 0.5f;
}
static v3 uv_to_direction(float u, float v)
{
    const float phi = v * FLX_PI_F, theta = (u * 2.0f - 1.0f) * FLX_PI_F;
    const float sinPhi = flx_sinf(phi), cosPhi = flx_cosf(phi), sinTh = flx_sinf(theta), cosTh = flx_cosf(theta);
    return V(sinPhi * sinTh, cosPhi, -sinPhi * cosTh);
}
static v3 eval_env_dir(const Scene *sc, v3 d) { float u, v, o[4]; direction_to_uv(d, &u, &v); flx_bilinear_rgba(sc->envRGBA, sc->envW, sc->envH, u, v, o); return V(o[0], o[1], o[2]); }
static void sample_env_alias(const Scene *sc, float r01, v3 *L, float *pdfW)
{
    const int width = sc->envW, height = sc->envH;
    const float r = r01 * width * height;
    int i = (int)floorf(r); if (i > width * height - 1) i = width * height - 1;
    const float mProb = sc->prob[i];
    const int uvInd = (r - i < mProb) ? i : sc->alias[i];
    const float pdf_uv = sc->pdf[uvInd];
    const int uInd = uvInd % width, vInd = uvInd / width;
    const float u = (float)(uInd + 0.5f) / width, v = (float)(vInd + 0.5f) / height;
    *L = uv_to_direction(u, v);
    const float sinTh = flx_sinf(FLX_PI_F * v);
    const float directPdfUV = pdf_uv * 1.0f;
    *pdfW = (sinTh != 0.0f) ? directPdfUV / (2.0f * FLX_PI_F * FLX_PI_F * sinTh) : 0.0f;
}
static float env_map_pdf(const Scene *sc, v3 d)
{
    float u, v; direction_to_uv(d, &u, &v);
    const float sinTh = flx_sinf(v * FLX_PI_F);
    if (sinTh == 0.0f) return 0.0f;
    int iu = (int)floorf(u * sc->envW); if (iu > sc->envW - 1) iu = sc->envW - 1;
    int iv = (int)floorf(v * sc->envH); if (iv > sc->envH - 1) iv = sc->envH - 1;
    return sc->pdf[iv * sc->envW + iu] / (FLX_2PI_F * FLX_PI_F * sinTh);
}

/* ================================================================ logic (wf_logic.cl:14-314) */
static float luminance(v3 v) { return 0.212671f * v.x + 0.715160f * v.y + 0.072169f * v.z; }
static float pdf_a_to_w(float pdf, float dist, float cosine) { return pdf * (dist * dist) / fabsf(cosine); }

static void logic_kernel(const RefBufs *b, size_t begin, size_t end, int separate)
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); QueueCounters *ql = (QueueCounters *)b->queueLens; const Scene sc = scene_of(b);
    const uint32_t wh = p->width * p->height;
    const uint32_t maxId = b->firstIteration ? (wh < b->numTasks ? wh : b->numTasks) : b->numTasks;
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= maxId) continue;
        uint32_t seed = ru(t, S_SEED, gid); const uint32_t len = ru(t, S_LEN, gid);
        Hit hit = read_hit(t, gid);
        const v3 rayOrig = rv(t, S_ORIG, gid), rayDir = rv(t, S_DIR, gid);
        v3 T = rv(t, S_T, gid);
        float contProb = 1.0f;
        int terminate = (len >= p->maxBounces + 1);
        if (terminate && p->useRoulette)
        {
            contProb = clampf(luminance(T), 0.01f, 0.5f);
            terminate = (rnd(&seed) > contProb);
            T = divs(T, contProb);
            wv(t, S_T, gid, T);
        }
        if (is_zero(T) || rf(t, S_LPDFW, gid) == 0.0f) terminate = 1;

        if (hit.i < 0 && !terminate)
        {
            float weight = 1.0f; const int lastSpecular = ru(t, S_LSPEC, gid) != 0; v3 bg = V1(0.0f);
            if (p->useEnvMap && (len == 1 || p->sampleImpl)) bg = scl(eval_env_dir(&sc, rayDir), p->envMapStrength);
            if (p->sampleImpl && p->sampleExpl && p->useEnvMap && len > 1 && !lastSpecular)
            {
                const float lightPickProb = rf(t, S_LPICK, gid), directPdfW = env_map_pdf(&sc, rayDir), actualPdfW = rf(t, S_LPDFW, gid);
                weight = (actualPdfW * lightPickProb) / (actualPdfW * lightPickProb + directPdfW);
            }
            wv(t, S_EI, gid, add(rv(t, S_EI, gid), mul(lscl(weight, T), bg)));
            terminate = 1;
        }
        else if (hit.areaLightHit && !terminate)
        {
            float misWeight = 1.0f; const int lastSpecular = ru(t, S_LSPEC, gid) != 0;
            if (p->sampleExpl && len > 1 && !lastSpecular)
            {
                const float directPdfA = 1.0f / (4.0f * p->areaLight.sizex * p->areaLight.sizey);
                const float directPdfW = pdf_a_to_w(directPdfA, length(sub(hit.P, rayOrig)), dot(normalize(neg(rayDir)), hit.N));
                const float lightPickProb = rf(t, S_LPICK, gid), lastPdfW = rf(t, S_LPDFW, gid);
                misWeight = lastPdfW / (lastPdfW + directPdfW * lightPickProb);
            }
            wv(t, S_EI, gid, add(rv(t, S_EI, gid), mul(scl(T, misWeight), F(p->areaLight.E))));
            terminate = 1;
        }

        if (!ru(t, S_BLOCKED, gid))
        {
            const v3 emission = rv(t, S_LEMIT, gid), bsdf = rv(t, S_LBSDF, gid);
            const float cosTh = rf(t, S_LCOSTH, gid), directPdfW = rf(t, S_LPDFDIRECT, gid), bsdfPdfW = rf(t, S_LPDFIMPL, gid), lightPickProb = rf(t, S_LPICK, gid);
            float weight = 1.0f;
            if (p->sampleImpl) weight = (directPdfW * lightPickProb) / (directPdfW * lightPickProb + bsdfPdfW);
            const v3 lastT = rv(t, S_LT, gid);
            const v3 contrib = divs(scl(scl(mul(mul(bsdf, lastT), emission), weight), cosTh), lightPickProb * directPdfW);
            wv(t, S_EI, gid, add(rv(t, S_EI, gid), contrib));
        }

        if (terminate)
        {
            if (len > 0)
            {
                const uint32_t pix = ru(t, S_PIXEL, gid); const v3 Ei = rv(t, S_EI, gid);
                add_float(&b->pixels[4 * (size_t)pix + 0], Ei.x); add_float(&b->pixels[4 * (size_t)pix + 1], Ei.y);
                add_float(&b->pixels[4 * (size_t)pix + 2], Ei.z); add_float(&b->pixels[4 * (size_t)pix + 3], 1.0f);
            }
            b->raygenQueue[atomic_inc(&ql->raygen)] = gid;
            wu(t, S_SEED, gid, seed);
            continue;
        }

        const Material mat = sc.materials[hit.matId];
        hit.N = tangent_space_normal(&hit, &mat, &sc);
        const int backface = dot(hit.N, rayDir) > 0.0f;
        if (backface) hit.N = scl(hit.N, -1.0f);
        const v3 orig = sub(hit.P, lscl(1e-3f, rayDir));
        write_hit(t, gid, &hit);
        wu(t, S_BACKFACE, gid, (uint32_t)backface);

        if (p->sampleExpl && !IS_SINGULAR(mat.type))
        {
            const uint32_t nl = p->useEnvMap + p->useAreaLight;
            const float envMapProb = (float)p->useEnvMap / (nl > 1u ? nl : 1u);
            const int useEnv = rnd(&seed) < envMapProb;
            const int useArea = !useEnv && p->useAreaLight;
            if (useEnv && p->useEnvMap)
            {
                v3 L; float directPdfW = 0.0f;
                sample_env_alias(&sc, rnd(&seed), &L, &directPdfW);
                const float lenL = 2.0f * p->worldRadius;
                L = normalize(L);
                const float cosTh = fmaxf(0.0f, dot(L, hit.N));
                const v3 Li = scl(eval_env_dir(&sc, L), p->envMapStrength);
                wv(t, S_SORIG, gid, orig); wv(t, S_SDIR, gid, L); wf(t, S_SLEN, gid, lenL); wf(t, S_LPDFDIRECT, gid, directPdfW); wf(t, S_LCOSTH, gid, cosTh);
                wf(t, S_LPICK, gid, envMapProb); wv(t, S_LEMIT, gid, Li);
                b->shadowQueue[atomic_inc(&ql->shadow)] = gid;
            }
            if (useArea)
            {
                const float lightPickProb = 1.0f - envMapProb;
                const AreaLight *A = &p->areaLight;
                const float directPdfA = 1.0f / (4.0f * A->sizex * A->sizey);                            /* sampleAreaLight, utils.cl:226-234 */
                v3 posL = F(A->pos);
                const float r1 = 2.0f * rnd(&seed) - 1.0f, r2 = 2.0f * rnd(&seed) - 1.0f;
                posL = add(posL, lscl(r1 * A->sizex, F(A->right))); posL = add(posL, lscl(r2 * A->sizey, F(A->up)));
                v3 L = sub(posL, orig);
                const float lenL = length(L) * 0.995f;
                L = normalize(L);
                const float cosLight = fmaxf(dot(F(A->N), neg(L)), 0.0f);
                if (cosLight > 0.0f)
                {
                    const float directPdfW = pdf_a_to_w(directPdfA, lenL, cosLight), cosTh = fmaxf(0.0f, dot(L, hit.N));
                    wv(t, S_SORIG, gid, orig); wv(t, S_SDIR, gid, L); wf(t, S_SLEN, gid, lenL); wf(t, S_LPDFDIRECT, gid, directPdfW); wf(t, S_LCOSTH, gid, cosTh);
                    wf(t, S_LPICK, gid, lightPickProb); wv(t, S_LEMIT, gid, F(A->E));
                    b->shadowQueue[atomic_inc(&ql->shadow)] = gid;
                }
                else wu(t, S_BLOCKED, gid, 1);
            }
        }
        wu(t, S_SEED, gid, seed);

        uint32_t *queue = b->diffuseQueue, *qlen = &ql->diffuse;                                         /* addToMaterialQueueNaive, wf_logic.cl:322-372 */
        if (separate)
        {
            switch (mat.type)
            {
            case BXDF_DIFFUSE: break;
            case BXDF_GLOSSY: queue = b->glossyQueue; qlen = &ql->glossy; break;
            case BXDF_GGX_REFL: queue = b->ggxReflQueue; qlen = &ql->ggxRefl; break;
            case BXDF_GGX_REFR: queue = b->ggxRefrQueue; qlen = &ql->ggxRefr; break;
            case BXDF_IDEAL_REFL: case BXDF_IDEAL_DIEL: queue = b->deltaQueue; qlen = &ql->delta; break;
            default: continue;
            }
        }
        queue[atomic_inc(qlen)] = gid;
    }
}
void NAME(logic_single)(const RefBufs *b, size_t begin, size_t end) { logic_kernel(b, begin, end, 0); }
void NAME(logic_separate)(const RefBufs *b, size_t begin, size_t end) { logic_kernel(b, begin, end, 1); }

/* ================================================================ BSDFs (diffuse.cl, ggx.cl, glossy.cl, ideal_*.cl, fresnel.cl) */
static v3 reflect(v3 d, v3 n) { return sub(d, lscl(2.0f * dot(d, n), n)); }
static v3 refract(v3 wi, v3 n, float eta)
{
    const float iDotN = dot(neg(wi), n), sin2I = fmaxf(0.0f, 1.0f - iDotN * iDotN), sin2T = eta * eta * sin2I, cosT = sqrtf(fmaxf(0.0f, 1.0f - sin2T));
    return add(scl(wi, eta), scl(n, eta * iDotN - cosT));
}
static void make_ortho_basis(v3 N, v3 *a, v3 *b)
{
    if (N.x != N.y || N.x != N.z) *a = V(N.z - N.y, N.x - N.z, N.y - N.x); else *a = V(N.z - N.y, N.x + N.z, -N.y - N.x);
    *a = normalize(*a); *b = cross(N, *a);
}
static v3 cos_sample_hemisphere(v3 n, uint32_t *seed, float *pdf)                                       /* utils.cl:82-112 */
{
    const float r1 = 2.0f * FLX_PI_F * rnd(seed), r2 = rnd(seed), r2s = sqrtf(r2);
    v3 w = n, u = (fabsf(w.x) > 0.1f) ? cross(V(0.0f, 1.0f, 0.0f), w) : cross(V(1.0f, 0.0f, 0.0f), w);
    u = normalize(u);
    v3 v = cross(w, u);
    u = scl(u, flx_cosf(r1) * r2s); v = scl(v, flx_sinf(r1) * r2s); w = scl(w, sqrtf(1 - r2));
    const v3 dir = add(add(u, v), w);
    *pdf = dot(n, dir) / FLX_PI_F;
    return dir;
}
static float fresnel_dielectric(float cosI, float etaI, float etaT)                                     /* fresnel.cl:5-20 */
{
    const float sinI = sqrtf(fmaxf(0.0f, 1.0f - cosI * cosI)), sinT = etaI / etaT * sinI, cosT = sqrtf(fmaxf(0.0f, 1.0f - sinT * sinT));
    if (sinT >= 1.0f) return 1.0f;
    const float parl = ((etaT * cosI) - (etaI * cosT)) / ((etaT * cosI) + (etaI * cosT)), perp = ((etaI * cosI) - (etaT * cosT)) / ((etaI * cosI) + (etaT * cosT));
    return 0.5f * (parl * parl + perp * perp);
}
static v3 eval_diffuse(const Hit *h, const Material *m, const Scene *sc) { return scl(mat_albedo(F(m->Kd), h->u, h->v, m->map_Kd, sc), FLX_INV_PI_F); }
static float pdf_diffuse(const Hit *h, v3 dirOut) { return dot(h->N, dirOut) * FLX_INV_PI_F; }
static v3 sample_diffuse(const Hit *h, const Material *m, const Scene *sc, v3 *dirOut, float *pdfW, uint32_t *seed)
{
    *dirOut = cos_sample_hemisphere(h->N, seed, pdfW);
    return eval_diffuse(h, m, sc);
}
static float to_roughness(float Ns) { return sqrtf(2.0f / (2.0f + Ns)); }
static v3 ggx_sample_lobe(float alpha, v3 N, uint32_t *seed)                                            /* ggx.cl:18-37 */
{
    v3 X, Y; make_ortho_basis(N, &X, &Y);
    const float r0 = rnd(seed), r1 = rnd(seed);
    const float theta = flx_atan2f(alpha * sqrtf(r0), sqrtf(1 - r0)), phi = FLX_2PI_F * r1;
    const float sinT = flx_sinf(theta), cosT = flx_cosf(theta), sinP = flx_sinf(phi), cosP = flx_cosf(phi);
    return normalize(add(add(scl(scl(X, sinT), cosP), scl(scl(Y, sinT), sinP)), scl(N, cosT)));
}
static float ggx_g1(float alpha, v3 v, v3 n, v3 m)
{
    const float mDotV = dot(m, v), nDotV = dot(n, v);
    if (nDotV * mDotV <= 0.0f) return 0.0f;
    const float c2 = nDotV * nDotV, tanSq = (c2 > 0.0f) ? ((1.0f - c2) / c2) : 0.0f;
    return 2.0f / (1.0f + sqrtf(1.0f + alpha * alpha * tanSq));
}
static float ggx_g(float alpha, v3 wi, v3 wo, v3 n, v3 m) { return ggx_g1(alpha, wi, n, m) * ggx_g1(alpha, wo, n, m); }
static float ggx_d(float alpha, v3 n, v3 m)
{
    const float nDotM = dot(n, m);
    if (nDotM <= 0.0f) return 0.0f;
    const float c2 = nDotM * nDotM, tanSq = nDotM != 0.0f ? ((1.0f - c2) / c2) : 0.0f, aSq = alpha * alpha;
    const float denom = FLX_PI_F * c2 * c2 * (aSq + tanSq) * (aSq + tanSq);
    return denom > 0.0f ? (aSq / denom) : 0.0f;
}
static float ggx_pdf_reflect(float alpha, v3 wo, v3 N, v3 H)
{
    const float nDotH = fabsf(dot(N, H)), oDotH = fabsf(dot(wo, H)), jInv = 4.0f * oDotH;
    return jInv == 0.0f ? 0.0f : ggx_d(alpha, N, H) * nDotH / jInv;
}
static v3 ggx_reflect_term(const Hit *h, const Material *m, const Scene *sc, v3 wi, v3 wo, v3 H)        /* ggx.cl:101-113 / 126-137 */
{
    const float alpha = to_roughness(m->Ns), iDotN = dot(wi, h->N), oDotN = dot(wo, h->N);
    const float Fr = (m->Ni > 1.0f) ? fresnel_dielectric(iDotN, 1.0f, m->Ni) : 1.0f;
    const v3 Ks = mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc);
    const float D = ggx_d(alpha, h->N, H), G = ggx_g(alpha, wi, wo, h->N, H), den = 4.0f * iDotN * oDotN;
    return (den != 0.0f) ? divs(scl(scl(scl(Ks, Fr), G), D), den) : V1(0.0f);
}
static v3 sample_ggx_reflect(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW, uint32_t *seed)
{
    const v3 wi = scl(dirIn, -1);
    const float alpha = to_roughness(m->Ns);
    const v3 H = ggx_sample_lobe(alpha, h->N, seed);
    *dirOut = reflect(neg(wi), H);
    *pdfW = ggx_pdf_reflect(alpha, *dirOut, h->N, H);
    return ggx_reflect_term(h, m, sc, wi, *dirOut, H);
}
static v3 eval_ggx_reflect(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 dirOut)
{
    const v3 wi = scl(dirIn, -1), H = normalize(add(wi, dirOut));
    return ggx_reflect_term(h, m, sc, wi, dirOut, H);
}
static float pdf_ggx_reflect(const Hit *h, const Material *m, v3 dirIn, v3 dirOut)
{
    const v3 wi = scl(dirIn, -1), H = normalize(add(wi, dirOut));
    return ggx_pdf_reflect(to_roughness(m->Ns), dirOut, h->N, H);
}
static float ggx_pdf_refract(float alpha, float etaI, float etaO, v3 wi, v3 wo, v3 N, v3 H)
{
    const float nDotH = fabsf(dot(N, H)), iDotH = fabsf(dot(wi, H)), oDotH = fabsf(dot(wo, H)), sj = etaI * iDotH + etaO * oDotH;
    return sj == 0.0f ? 0.0f : ggx_d(alpha, N, H) * nDotH * oDotH * etaO * etaO / (sj * sj);
}
static v3 ggx_transmit_term(const Hit *h, const Material *m, const Scene *sc, float alpha, float etaI, float etaO, float Fr, float iDotN, float oDotN,
                            float iDotH, float oDotH, v3 wi, v3 wo, v3 Nn, v3 H)                         /* ggx.cl:196-218 / 252-271 */
{
    const float eta = etaI / etaO;
    v3 bsdf = V1(eta * eta);
    bsdf = mul(bsdf, mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc));
    const float denom = iDotN * oDotN * (etaI * iDotH + etaO * oDotH) * (etaI * iDotH + etaO * oDotH);
    if (denom == 0.0f) return V1(0.0f);
    const float focus = etaO * etaO * iDotH * oDotH / denom, D = ggx_d(alpha, Nn, H), G = ggx_g(alpha, wi, wo, Nn, H);
    return scl(scl(scl(lscl(1.0f - Fr, bsdf), D), G), focus);
}
static v3 sample_ggx_refract(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW, uint32_t *seed)
{
    const v3 wi = scl(dirIn, -1);
    const float raylen = length(wi), alpha = to_roughness(m->Ns);
    float etaI = 1.0f, etaO = m->Ni;
    if (backface) { const float tmp = etaI; etaI = etaO; etaO = tmp; }
    const float iDotN = dot(normalize(wi), h->N);
    v3 H = ggx_sample_lobe(alpha, h->N, seed);
    const float Fr = fresnel_dielectric(iDotN, etaI, etaO);
    if (rnd(seed) < Fr)
    {
        *dirOut = lscl(raylen, reflect(normalize(neg(wi)), H));
        *pdfW = ggx_pdf_reflect(alpha, *dirOut, h->N, H);
        const float oDotN = dot(*dirOut, h->N), D = ggx_d(alpha, h->N, H), G = ggx_g(alpha, wi, *dirOut, h->N, H), den = 4.0f * iDotN * oDotN;
        return V1((den != 0.0f) ? (Fr * G * D / den) : 0.0f);
    }
    const float eta = etaI / etaO;
    *dirOut = lscl(raylen, refract(normalize(neg(wi)), h->N, eta));
    H = normalize(neg(add(scl(wi, etaI), scl(*dirOut, etaO))));
    const v3 Nn = backface ? neg(h->N) : h->N;
    *pdfW = ggx_pdf_refract(alpha, etaI, etaO, wi, *dirOut, Nn, H);
    const float iDotH = fabsf(dot(normalize(wi), H)), oDotH = fabsf(dot(*dirOut, H)), oDotN = dot(*dirOut, h->N);
    return ggx_transmit_term(h, m, sc, alpha, etaI, etaO, Fr, iDotN, oDotN, iDotH, oDotH, wi, *dirOut, Nn, H);
}
static v3 eval_ggx_refract(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 dirOut)
{
    const v3 wi = scl(dirIn, -1);
    const float alpha = to_roughness(m->Ns);
    float etaI = 1.0f, etaO = m->Ni;
    if (backface) { const float tmp = etaI; etaI = etaO; etaO = tmp; }
    const float iDotN = dot(normalize(wi), h->N), oDotN = dot(normalize(dirOut), h->N), Fr = fresnel_dielectric(iDotN, etaI, etaO);
    if (!backface)
    {
        const v3 H = normalize(add(wi, dirOut));
        const float D = ggx_d(alpha, h->N, H), G = ggx_g(alpha, wi, dirOut, h->N, H), den = 4.0f * iDotN * oDotN;
        return (den != 0.0f) ? V1(Fr * G * D / den) : V1(0.0f);
    }
    const v3 H = normalize(neg(add(scl(wi, etaI), scl(dirOut, etaO))));
    const float iDotH = fabsf(dot(normalize(wi), H)), oDotH = fabsf(dot(normalize(dirOut), H));
    return ggx_transmit_term(h, m, sc, alpha, etaI, etaO, Fr, iDotN, oDotN, iDotH, oDotH, wi, dirOut, neg(h->N), H);
}
static float pdf_ggx_refract(const Hit *h, const Material *m, int backface, v3 dirIn, v3 dirOut)
{
    const v3 wi = scl(dirIn, -1);
    const float alpha = to_roughness(m->Ns);
    if (!backface) { const v3 H = normalize(add(wi, dirOut)); return ggx_pdf_reflect(alpha, dirOut, h->N, H); }
    const float etaI = m->Ni, etaO = 1.0f;
    const v3 H = normalize(neg(add(scl(wi, etaI), scl(dirOut, etaO))));
    return ggx_pdf_refract(alpha, etaI, etaO, wi, dirOut, neg(h->N), H);
}
static float ks_to_eta(v3 Ks) { const float k = clampf((Ks.x + Ks.y + Ks.z) / 3.0f, 0.0f, 0.99f); return (sqrtf(k) + 1) / (1 - sqrtf(k)); } /* glossy.cl:18-22 */
static v3 eta_to_ks(float eta) { const float r = (eta > 0.0f) ? ((eta - 1) / (eta + 1)) : 0.0f; return V1(r * r); }
static Material glossy_material(const Hit *h, const Material *m, const Scene *sc, int by_length)
{
    Material e = *m;
    const v3 Ks = mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc);
    e.Ks.x = Ks.x; e.Ks.y = Ks.y; e.Ks.z = Ks.z;
    e.Ni = (m->Ni > 0.0f) ? m->Ni : ks_to_eta(Ks);
    if (by_length ? (length(Ks) == 0.0f) : is_zero(Ks)) { const v3 k = eta_to_ks(e.Ni); e.Ks.x = k.x; e.Ks.y = k.y; e.Ks.z = k.z; }
    return e;
}
static v3 sample_glossy(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW, uint32_t *seed)   /* glossy.cl:24-62 */
{
    const Material e = glossy_material(h, m, sc, 0);
    const float cosTh = dot(normalize(neg(dirIn)), h->N), Fr = fresnel_dielectric(cosTh, 1.0f, e.Ni);
    float basePdf, coatPdf; v3 base, coat;
    if (rnd(seed) < Fr) { coat = sample_ggx_reflect(h, &e, sc, dirIn, dirOut, &coatPdf, seed); base = eval_diffuse(h, &e, sc); basePdf = pdf_diffuse(h, *dirOut); }
    else { base = sample_diffuse(h, &e, sc, dirOut, &basePdf, seed); coat = eval_ggx_reflect(h, &e, sc, dirIn, *dirOut); coatPdf = pdf_ggx_reflect(h, &e, dirIn, *dirOut); }
    if (dot(h->N, *dirOut) < 1e-5f) return V1(0.0f);   /* pdfW stays as the caller left it (unset in the reference) */
    *pdfW = (1 - Fr) * basePdf + Fr * coatPdf;
    return add(scl(base, 1 - Fr), coat);
}
static v3 eval_glossy(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 dirOut)
{
    const Material e = glossy_material(h, m, sc, 1);
    const v3 base = eval_diffuse(h, &e, sc), coat = eval_ggx_reflect(h, &e, sc, dirIn, dirOut);
    const float cosTh = dot(normalize(neg(dirIn)), h->N), Fr = fresnel_dielectric(cosTh, 1.0f, e.Ni);
    return add(scl(base, 1 - Fr), coat);
}
static float pdf_glossy(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 dirOut)
{
    const v3 Ks = mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc);
    const float Ni = (m->Ni > 0.0f) ? m->Ni : ks_to_eta(Ks);
    const float basePdf = pdf_diffuse(h, dirOut), coatPdf = pdf_ggx_reflect(h, m, dirIn, dirOut);
    const float cosTh = dot(normalize(neg(dirIn)), h->N), Fr = fresnel_dielectric(cosTh, 1.0f, Ni);
    return (1 - Fr) * basePdf + Fr * coatPdf;
}
static v3 sample_ideal_reflection(const Hit *h, const Material *m, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW)        /* ideal_reflection.cl:9-22 */
{
    const float len = length(dirIn);
    *dirOut = lscl(len, reflect(normalize(dirIn), h->N)); *pdfW = 1.0f;
    const v3 ks = mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc);
    const float cosO = dot(normalize(*dirOut), h->N);
    return (cosO != 0.0f) ? divs(ks, cosO) : V1(0.0f);
}
static v3 sample_ideal_dielectric(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW, uint32_t *seed) /* ideal_dielectric.cl:10-44 */
{
    const float raylen = length(dirIn);
    v3 bsdf = V1(1.0f);
    const float cosI = dot(normalize(neg(dirIn)), h->N);
    float n1 = 1.0f, n2 = m->Ni;
    if (backface) { const float tmp = n1; n1 = n2; n2 = tmp; }
    const float eta = n1 / n2, fr = fresnel_dielectric(cosI, n1, n2);
    if (rnd(seed) < fr) *dirOut = lscl(raylen, reflect(normalize(dirIn), h->N));
    else { *dirOut = lscl(raylen, refract(normalize(dirIn), h->N, eta)); bsdf = scl(bsdf, eta * eta); bsdf = mul(bsdf, mat_float3(F(m->Ks), h->u, h->v, m->map_Ks, sc)); }
    *pdfW = 1.0f;
    const float cosO = dot(normalize(*dirOut), h->N);
    return divs(bsdf, cosO);
}

/* type dispatch (bxdf_partial.cl:19-153); `mask` = which lobes the kernel was compiled with */
static v3 bxdf_eval(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 dirOut, int mask)
{
    switch (m->type & mask)
    {
    case BXDF_DIFFUSE: return eval_diffuse(h, m, sc);
    case BXDF_GLOSSY: return eval_glossy(h, m, sc, dirIn, dirOut);
    case BXDF_GGX_REFL: return eval_ggx_reflect(h, m, sc, dirIn, dirOut);
    case BXDF_GGX_REFR: return eval_ggx_refract(h, m, backface, sc, dirIn, dirOut);
    case BXDF_EMISSIVE: return V1(1.0f);
    }
    return V1(0.0f);
}
static float bxdf_pdf(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 dirOut, int mask)
{
    switch (m->type & mask)
    {
    case BXDF_DIFFUSE: return pdf_diffuse(h, dirOut);
    case BXDF_GLOSSY: return pdf_glossy(h, m, sc, dirIn, dirOut);
    case BXDF_GGX_REFL: return pdf_ggx_reflect(h, m, dirIn, dirOut);
    case BXDF_GGX_REFR: return pdf_ggx_refract(h, m, backface, dirIn, dirOut);
    }
    return 0.0f;
}
static v3 bxdf_sample(const Hit *h, const Material *m, int backface, const Scene *sc, v3 dirIn, v3 *dirOut, float *pdfW, uint32_t *seed, int mask)
{
    switch (m->type & mask)
    {
    case BXDF_DIFFUSE: return sample_diffuse(h, m, sc, dirOut, pdfW, seed);
    case BXDF_GLOSSY: return sample_glossy(h, m, sc, dirIn, dirOut, pdfW, seed);
    case BXDF_GGX_REFL: return sample_ggx_reflect(h, m, sc, dirIn, dirOut, pdfW, seed);
    case BXDF_IDEAL_REFL: return sample_ideal_reflection(h, m, sc, dirIn, dirOut, pdfW);
    case BXDF_GGX_REFR: return sample_ggx_refract(h, m, backface, sc, dirIn, dirOut, pdfW, seed);
    case BXDF_IDEAL_DIEL: return sample_ideal_dielectric(h, m, backface, sc, dirIn, dirOut, pdfW, seed);
    case BXDF_EMISSIVE: return V1(1.0f);
    }
    return V1(0.0f);
}

/* ================================================================ material kernels (wf_mat_*.cl) */
static void material_kernel(const RefBufs *b, size_t begin, size_t end, const uint32_t *queue, const uint32_t *qlen, int mask)
{
    Tasks t = tasks_of(b); QueueCounters *ql = (QueueCounters *)b->queueLens; const Scene sc = scene_of(b);
    LOOP
    {
        const uint32_t gd = (uint32_t)g_;
        if (gd >= *qlen) continue;
        const uint32_t gid = queue[gd];
        uint32_t seed = ru(t, S_SEED, gid);
        const Hit hit = read_hit(t, gid);
        const Material mat = sc.materials[hit.matId];
        const int backface = ru(t, S_BACKFACE, gid) != 0;
        const v3 dirIn = rv(t, S_DIR, gid), L = rv(t, S_SDIR, gid);
        const v3 bsdfNEE = bxdf_eval(&hit, &mat, backface, &sc, dirIn, L, mask);
        const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf(&hit, &mat, backface, &sc, dirIn, L, mask));
        wv(t, S_LBSDF, gid, bsdfNEE); wf(t, S_LPDFIMPL, gid, bsdfPdfW);
        float pdfW = 0.0f; v3 newDir = V1(0.0f);   /* the reference leaves both uninitialised; zero is this repo's pinned choice */
        const v3 bsdf = bxdf_sample(&hit, &mat, backface, &sc, dirIn, &newDir, &pdfW, &seed, mask);
        const float costh = dot(hit.N, normalize(newDir));
        const v3 oldT = rv(t, S_T, gid);
        v3 newT = V1(0.0f);
        if (!(pdfW == 0.0f || is_zero(bsdf))) newT = divs(scl(mul(oldT, bsdf), costh), pdfW);
        const v3 orig = add(hit.P, lscl(1e-4f, newDir));
        wv(t, S_LT, gid, oldT); wv(t, S_T, gid, newT); wv(t, S_ORIG, gid, orig); wv(t, S_DIR, gid, newDir); wf(t, S_LPDFW, gid, pdfW);
        wu(t, S_SEED, gid, seed); wu(t, S_LSPEC, gid, IS_SINGULAR(mat.type) ? 1u : 0u);
        b->extensionQueue[atomic_inc(&ql->extension)] = gid;
    }
}
#define QL(b) ((QueueCounters *)(b)->queueLens)
void NAME(mat_all)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->diffuseQueue, &QL(b)->diffuse, 0xfe); }
void NAME(mat_diffuse)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->diffuseQueue, &QL(b)->diffuse, BXDF_DIFFUSE); }
void NAME(mat_glossy)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->glossyQueue, &QL(b)->glossy, BXDF_GLOSSY); }
void NAME(mat_ggx_refl)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->ggxReflQueue, &QL(b)->ggxRefl, BXDF_GGX_REFL); }
void NAME(mat_ggx_refr)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->ggxRefrQueue, &QL(b)->ggxRefr, BXDF_GGX_REFR); }
void NAME(mat_delta)(const RefBufs *b, size_t begin, size_t end) { material_kernel(b, begin, end, b->deltaQueue, &QL(b)->delta, BXDF_IDEAL_REFL | BXDF_IDEAL_DIEL); }

/* ================================================================ microkernel integrator (src/mk_*.cl)
 * The reference's other integrator: one path per pixel, a phase word per path (geom.h:184-193).  Launch shapes as in
 * clcontext.cpp:709-750: reset / splat / splatPreview over width*height, the rest over numTasks; all clamp to
 * limit = min(width * height, numTasks).  The 2-D kernels compute gid = x + y * width, i.e. the flattened index. */
enum { MK_RT_NEXT_VERTEX = 0, MK_SAMPLE_BSDF = 1, MK_SPLAT_SAMPLE = 4, MK_GENERATE_CAMERA_RAY = 5 };
#define S_PHASE 46
static inline uint32_t mk_limit(const RefBufs *b) { const RenderParams *p = (const RenderParams *)b->params; const uint32_t n = p->width * p->height; return n < b->numTasks ? n : b->numTasks; }

void NAME(mk_reset)(const RefBufs *b, size_t begin, size_t end)                                         /* mk_reset.cl:4-43 */
{
    Tasks t = tasks_of(b); const uint32_t limit = mk_limit(b);
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        for (int c = 0; c < 4; c++) { b->pixels[4 * (size_t)gid + c] = 0.0f; b->denoiserNormal[4 * (size_t)gid + c] = 0.0f; }
        b->denoiserAlbedo[4 * (size_t)gid + 0] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 1] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 2] = 0.1f; b->denoiserAlbedo[4 * (size_t)gid + 3] = 0.0f;
        wu(t, S_PHASE, gid, MK_GENERATE_CAMERA_RAY);
        wv(t, S_EI, gid, V1(0.0f)); wv(t, S_T, gid, V1(1.0f)); wu(t, S_LEN, gid, 0); wu(t, S_LSPEC, gid, 1); wf(t, S_LPDFW, gid, 1.0f);
        wu(t, S_FIRSTDIFF, gid, 0); wu(t, S_SEED, gid, gid);
    }
}

void NAME(mk_raygen)(const RefBufs *b, size_t begin, size_t end)                                        /* mk_raygen.cl:5-63 */
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); const uint32_t limit = mk_limit(b);
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        if (ru(t, S_PHASE, gid) != MK_GENERATE_CAMERA_RAY) continue;
        uint32_t seed = ru(t, S_SEED, gid);
        float x = (float)(gid % p->width), y = (float)(gid / p->width);
        x += rnd(&seed); y += rnd(&seed);
        const float NDCx = x / p->width, NDCy = y / p->height;
        float SCRx = 2.0f * NDCx - 1.0f, SCRy = 2.0f * NDCy - 1.0f;
        SCRx *= (float)p->width / p->height;
        const float scale = flx_tanf(0.5f * p->camera.fov * 3.14159265358979323846f / 180);
        SCRx *= scale; SCRy *= scale;
        v3 rayOrig = F(p->camera.pos);
        const v3 target = add(add(add(rayOrig, scl(F(p->camera.right), SCRx)), scl(F(p->camera.up), SCRy)), F(p->camera.dir));
        v3 rayDir = normalize(sub(target, rayOrig));
        const v3 fp = add(F(p->camera.pos), scl(rayDir, p->camera.focalDist));
        const float sqrt_r = sqrtf(rnd(&seed)); const float th = FLX_2PI_F * rnd(&seed);
        const float rx = sqrt_r * flx_cosf(th), ry = sqrt_r * flx_sinf(th);
        rayOrig = add(rayOrig, lscl(p->worldRadius * p->camera.apertureSize, add(scl(F(p->camera.right), rx), scl(F(p->camera.up), ry))));
        rayDir = normalize(sub(fp, rayOrig));
        wv(t, S_ORIG, gid, rayOrig); wv(t, S_DIR, gid, rayDir);
        wu(t, S_SEED, gid, seed); wu(t, S_PHASE, gid, MK_RT_NEXT_VERTEX);
    }
}

void NAME(mk_next_vertex)(const RefBufs *b, size_t begin, size_t end)                                   /* mk_next_vertex.cl:7-123 */
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); const Scene sc = scene_of(b); const uint32_t limit = mk_limit(b);
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        if (ru(t, S_PHASE, gid) != MK_RT_NEXT_VERTEX) continue;
        const v3 o = rv(t, S_ORIG, gid), d = rv(t, S_DIR, gid);
        Hit hit = empty_hit(FLT_MAX);
        bvh_intersect(o, d, &hit, (const Triangle *)b->tris, (const Node *)b->nodes, b->indices, 0);
        if (p->sampleImpl && p->useAreaLight) intersect_light(&hit, o, d, p);
        write_hit(t, gid, &hit);
        uint32_t len = ru(t, S_LEN, gid);
        atomic_inc(len == 0 ? &b->stats[0] : &b->stats[1]);                                               /* primaryRays : extensionRays */
        len += 1; wu(t, S_LEN, gid, len);
        if (hit.i < 0)                                                                                  /* implicit environment sample */
        {
            v3 bg = V1(0.0f);
            if (p->useEnvMap && (len == 1 || p->sampleImpl)) bg = scl(eval_env_dir(&sc, d), p->envMapStrength);
            float weight = 1.0f;
            const int lastSpecular = ru(t, S_LSPEC, gid) != 0;
            if (p->sampleImpl && p->sampleExpl && p->useEnvMap && len > 1 && !lastSpecular)
            {
                const float lightPickProb = 1.0f;
                const float directPdfW = env_map_pdf(&sc, d), actualPdfW = rf(t, S_LPDFW, gid);
                weight = (actualPdfW * lightPickProb) / (actualPdfW * lightPickProb + directPdfW);
            }
            wv(t, S_EI, gid, add(rv(t, S_EI, gid), mul(lscl(weight, rv(t, S_T, gid)), bg)));
            wu(t, S_PHASE, gid, MK_SPLAT_SAMPLE);
        }
        else if (hit.areaLightHit)                                                                      /* implicit area-light sample */
        {
            float misWeight = 1.0f;
            const int lastSpecular = ru(t, S_LSPEC, gid) != 0;
            if (p->sampleExpl && len > 1 && !lastSpecular)
            {
                const float directPdfA = 1.0f / (4.0f * p->areaLight.sizex * p->areaLight.sizey);
                const float directPdfW = pdf_a_to_w(directPdfA, length(sub(hit.P, o)), dot(normalize(neg(d)), hit.N));
                const float lightPickProb = 1.0f, lastPdfW = rf(t, S_LPDFW, gid);
                misWeight = lastPdfW / (lastPdfW + directPdfW * lightPickProb);
            }
            wv(t, S_EI, gid, add(rv(t, S_EI, gid), mul(scl(rv(t, S_T, gid), misWeight), F(p->areaLight.E))));
            wu(t, S_PHASE, gid, MK_SPLAT_SAMPLE);
        }
        else
            wu(t, S_PHASE, gid, MK_SAMPLE_BSDF);
    }
}

void NAME(mk_sample_bsdf)(const RefBufs *b, size_t begin, size_t end)                                   /* mk_sample_bsdf.cl:11-197 */
{
    const RenderParams *p = (const RenderParams *)b->params; Tasks t = tasks_of(b); const Scene sc = scene_of(b); const uint32_t limit = mk_limit(b);
    const Triangle *tris = (const Triangle *)b->tris; const Node *nodes = (const Node *)b->nodes;
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        uint32_t seed = ru(t, S_SEED, gid);
        if (ru(t, S_PHASE, gid) != MK_SAMPLE_BSDF) continue;
        const v3 rayDir = rv(t, S_DIR, gid);
        Hit hit = read_hit(t, gid);
        const Material mat = sc.materials[hit.matId];
        hit.N = tangent_space_normal(&hit, &mat, &sc);
        const int backface = dot(hit.N, rayDir) > 0.0f;
        if (backface) hit.N = scl(hit.N, -1.0f);
        v3 orig = sub(hit.P, lscl(1e-3f, rayDir));
        if (p->sampleExpl && !IS_SINGULAR(mat.type))                                                    /* next-event estimation */
        {
            const float lightPickProb = 1.0f;
            if (p->useEnvMap)
            {
                v3 L; float directPdfW = 0.0f;
                sample_env_alias(&sc, rnd(&seed), &L, &directPdfW);
                const float lenL = 2.0f * p->worldRadius;
                L = normalize(L);
                Hit hitL = empty_hit(lenL);
                if (p->useAreaLight) intersect_light(&hitL, orig, L, p);
                const int occluded = (hitL.i > -1) || bvh_occluded(orig, L, lenL, tris, nodes, b->indices, 0);
                atomic_inc(&b->stats[2]);
                if (!occluded && directPdfW != 0.0f)
                {
                    const v3 brdf = bxdf_eval(&hit, &mat, backface, &sc, rayDir, L, 0xfe);
                    const float cosTh = fmaxf(0.0f, dot(L, hit.N));
                    const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf(&hit, &mat, backface, &sc, rayDir, L, 0xfe));
                    float weight = 1.0f;
                    if (p->sampleImpl) weight = (directPdfW * lightPickProb) / (directPdfW * lightPickProb + bsdfPdfW);
                    const v3 T = rv(t, S_T, gid);
                    const v3 envMapLi = scl(eval_env_dir(&sc, L), p->envMapStrength);
                    const v3 contrib = divs(scl(scl(mul(mul(brdf, T), envMapLi), weight), cosTh), lightPickProb * directPdfW);
                    wv(t, S_EI, gid, add(rv(t, S_EI, gid), contrib));
                }
            }
            if (p->useAreaLight)
            {
                const AreaLight *A = &p->areaLight;
                const float directPdfA = 1.0f / (4.0f * A->sizex * A->sizey);                           /* sampleAreaLight, utils.cl:226-234 */
                v3 posL = F(A->pos);
                const float r1 = 2.0f * rnd(&seed) - 1.0f, r2 = 2.0f * rnd(&seed) - 1.0f;
                posL = add(posL, lscl(r1 * A->sizex, F(A->right)));
                posL = add(posL, lscl(r2 * A->sizey, F(A->up)));
                v3 L = sub(posL, orig);
                const float lenL = length(L);
                L = normalize(L);
                const int occluded = bvh_occluded(orig, L, lenL, tris, nodes, b->indices, 0);
                atomic_inc(&b->stats[2]);
                const float cosLight = fmaxf(dot(F(A->N), neg(L)), 0.0f);
                if (!occluded && cosLight > 0.0f)
                {
                    const v3 brdf = bxdf_eval(&hit, &mat, backface, &sc, rayDir, L, 0xfe);
                    const float cosTh = fmaxf(0.0f, dot(L, hit.N));
                    const float directPdfW = pdf_a_to_w(directPdfA, lenL, cosLight);
                    const float bsdfPdfW = fmaxf(0.0f, bxdf_pdf(&hit, &mat, backface, &sc, rayDir, L, 0xfe));
                    float weight = 1.0f;
                    if (p->sampleImpl) weight = (directPdfW * lightPickProb) / (directPdfW * lightPickProb + bsdfPdfW);
                    const v3 T = rv(t, S_T, gid);
                    const v3 contrib = divs(scl(scl(mul(mul(brdf, T), F(A->E)), weight), cosTh), lightPickProb * directPdfW);
                    wv(t, S_EI, gid, add(rv(t, S_EI, gid), contrib));
                }
            }
        }
        float contProb = 1.0f;                                                                          /* Russian roulette */
        const uint32_t len = ru(t, S_LEN, gid);
        int terminate = (len - 1 >= p->maxBounces);
        if (terminate && p->useRoulette)
        {
            contProb = clampf(luminance(rv(t, S_T, gid)), 0.01f, 0.5f);
            terminate = (rnd(&seed) > contProb);
        }
        float pdfW = 0.0f; v3 newDir = V1(0.0f);   /* the reference leaves both uninitialised; zero is this repo's pinned choice */
        const v3 bsdf = bxdf_sample(&hit, &mat, backface, &sc, rayDir, &newDir, &pdfW, &seed, 0xfe);
        const float costh = dot(hit.N, normalize(newDir));
        pdfW *= contProb;
        if (pdfW == 0.0f || is_zero(bsdf)) terminate = 1;
        const v3 newT = divs(scl(mul(rv(t, S_T, gid), bsdf), costh), pdfW);
        orig = add(hit.P, lscl(1e-4f, newDir));
        wv(t, S_T, gid, newT); wv(t, S_ORIG, gid, orig); wv(t, S_DIR, gid, newDir); wf(t, S_LPDFW, gid, pdfW); wu(t, S_SEED, gid, seed);
        wu(t, S_LSPEC, gid, IS_SINGULAR(mat.type) ? 1u : 0u);
        wu(t, S_PHASE, gid, terminate ? MK_SPLAT_SAMPLE : MK_RT_NEXT_VERTEX);
    }
}

void NAME(mk_splat)(const RefBufs *b, size_t begin, size_t end)                                         /* mk_splat.cl:5-41 */
{
    Tasks t = tasks_of(b); const uint32_t limit = mk_limit(b);
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        if (ru(t, S_PHASE, gid) != MK_SPLAT_SAMPLE) continue;
        const v3 Ei = rv(t, S_EI, gid);
        float color[4] = {Ei.x, Ei.y, Ei.z, 1.0f};
        float *px = b->pixels + 4 * (size_t)gid;
        if (px[3] > 0.0f) for (int c = 0; c < 4; c++) color[c] += px[c];
        for (int c = 0; c < 4; c++) px[c] = color[c];
        atomic_inc(&b->stats[3]);
        wv(t, S_EI, gid, V1(0.0f)); wv(t, S_T, gid, V1(1.0f)); wu(t, S_LEN, gid, 0); wu(t, S_FIRSTDIFF, gid, 0);
        wu(t, S_PHASE, gid, MK_GENERATE_CAMERA_RAY);
    }
}

void NAME(mk_splat_preview)(const RefBufs *b, size_t begin, size_t end)                                 /* mk_splat_preview.cl:5-25 */
{
    Tasks t = tasks_of(b); const uint32_t limit = mk_limit(b);
    LOOP
    {
        const uint32_t gid = (uint32_t)g_;
        if (gid >= limit) continue;
        const v3 Ei = rv(t, S_EI, gid);
        float *px = b->pixels + 4 * (size_t)gid;
        px[0] = Ei.x; px[1] = Ei.y; px[2] = Ei.z; px[3] = 0.0f;
        wv(t, S_EI, gid, V1(0.0f)); wv(t, S_T, gid, V1(1.0f)); wu(t, S_LEN, gid, 0);
        wu(t, S_PHASE, gid, MK_GENERATE_CAMERA_RAY);
    }
}

/* ================================================================ post-process (mk_postprocess.cl:7-55, tonemap.cl:3-26) */
static v3 uc2_tonemap_func(v3 x)
{
    const float A = 0.22, B = 0.30, C = 0.10, D = 0.20, E = 0.01, Fq = 0.30; /* double literals rounded to float, as in the reference */
    const v3 num = V(x.x * (A * x.x + C * B) + D * E, x.y * (A * x.y + C * B) + D * E, x.z * (A * x.z + C * B) + D * E);
    const v3 den = V(x.x * (A * x.x + B) + D * Fq, x.y * (A * x.y + B) + D * Fq, x.z * (A * x.z + B) + D * Fq);
    return V(num.x / den.x - E / Fq, num.y / den.y - E / Fq, num.z / den.z - E / Fq);
}
void NAME(postprocess)(const RefBufs *b, size_t begin, size_t end)
{
    const RenderParams *p = (const RenderParams *)b->params;
    const uint32_t limit = p->width * p->height;
    LOOP
    {
        const size_t gid = (size_t)g_;
        if (gid >= limit) continue;
        float c[4] = {b->pixels[4 * gid], b->pixels[4 * gid + 1], b->pixels[4 * gid + 2], b->pixels[4 * gid + 3]};
        if (c[3] > 0.0) { const float w = c[3]; c[0] /= w; c[1] /= w; c[2] /= w; c[3] /= w; }
        v3 col = V(c[0] * p->exposure, c[1] * p->exposure, c[2] * p->exposure);
        if (p->tmOperator == 1) col = V(col.x / (1.0f + col.x), col.y / (1.0f + col.y), col.z / (1.0f + col.z));
        if (p->tmOperator == 2)
        {
            const float W = 11.2, exposureBias = 2.0;
            const v3 a = uc2_tonemap_func(lscl(exposureBias, col)), w = uc2_tonemap_func(V(W, W, W));
            col = V(a.x / w.x, a.y / w.y, a.z / w.z);
        }
        col = V(flx_powf(col.x, 1.0f / 2.2f), flx_powf(col.y, 1.0f / 2.2f), flx_powf(col.z, 1.0f / 2.2f));
        b->pixelsPreview[4 * gid] = col.x; b->pixelsPreview[4 * gid + 1] = col.y; b->pixelsPreview[4 * gid + 2] = col.z; b->pixelsPreview[4 * gid + 3] = c[3];
    }
}

#ifndef PORT_PARALLEL
/* vector entry points for tests/test_math.py */
void port_math(int fn, const float *a, const float *bb, float *out, int n)
{
    for (int i = 0; i < n; i++)
        switch (fn)
        {
        case 0: out[i] = flx_sinf(a[i]); break;
        case 1: out[i] = flx_cosf(a[i]); break;
        case 2: out[i] = flx_tanf(a[i]); break;
        case 3: out[i] = flx_acosf(a[i]); break;
        case 4: out[i] = flx_atan2f(a[i], bb[i]); break;
        case 5: out[i] = flx_powf(a[i], bb[i]); break;
        }
}
void port_rand(uint32_t seed, float *out, uint32_t *seeds, int n) { for (int i = 0; i < n; i++) { out[i] = rnd(&seed); seeds[i] = seed; } }
#endif
