/*
 * bvh_oracle.c -- TEST INFRASTRUCTURE: CPU restatement of the GPU hierarchy builder (fluctus_b200/csrc/flx_bvh_build.cuh).
 *
 * Only tests/ load this (through oracle/liboracle.so); the product never does.  The builder is this repo's own
 * algorithm (the reference builds SBVHs on the CPU, src/sbvh.cpp), so there is no reference vector to pin it to: the pin
 * is (a) this independent, sequential, recursive statement of the same definition -- sort unique 62-bit keys, split every
 * range at the highest differing key bit, fit boxes and SAH costs bottom-up, collapse, emit depth first -- which the GPU
 * output must match bit for bit, and (b) the structural checks and the render parity against the reference's SBVH in
 * tests/ (same closest hits).
 *
 * Output format = the reference's (src/bvhnode.hpp:50-59; flattening order src/sbvh.cpp:52-73).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef struct { float x, y, z, w; } f4;
typedef struct { f4 bmin, bmax; int32_t parent; uint32_t link; uint8_t nPrims; uint8_t pad[7]; } Node;
typedef struct { f4 p, n, t; } Vertex;
typedef struct { Vertex v0, v1, v2; int32_t matId; int32_t pad[3]; } Triangle;

typedef struct
{
    const uint64_t *keys; const f4 *pmin, *pmax; uint32_t maxLeaf;
    Node *out; uint32_t nOut;
} Ctx;

static uint32_t spread10(uint32_t v)
{
    v = (v | (v << 16)) & 0x030000ffu; v = (v | (v << 8)) & 0x0300f00fu; v = (v | (v << 4)) & 0x030c30c3u; v = (v | (v << 2)) & 0x09249249u;
    return v;
}
static uint32_t quantize10(float c, float lo, float hi)
{
    const float ext = hi - lo;
    if (!(ext > 0.0f)) return 0u;
    const float q = ((c - lo) / ext) * 1024.0f;
    return (uint32_t)fminf(fmaxf(q, 0.0f), 1023.0f);
}
static int cmp_u64(const void *a, const void *b) { const uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b; return x < y ? -1 : (x > y ? 1 : 0); }
static float half_area(f4 lo, f4 hi) { const float dx = hi.x - lo.x, dy = hi.y - lo.y, dz = hi.z - lo.z; return (dx * dy + dy * dz) + dz * dx; }

/* a subtree before emission */
typedef struct Sub { f4 lo, hi; float cost; uint32_t size, first, last; int leaf; struct Sub *l, *r; } Sub;

static Sub *build(const Ctx *c, uint32_t first, uint32_t last)
{
    Sub *s = (Sub *)calloc(1, sizeof(Sub));
    s->first = first; s->last = last;
    if (first == last)
    {
        const uint32_t tri = (uint32_t)(c->keys[first] & 0xffffffffu);
        s->lo = c->pmin[tri]; s->hi = c->pmax[tri]; s->cost = half_area(s->lo, s->hi) * 1.0f; s->size = 1; s->leaf = 1;
        return s;
    }
    /* split after the last key that shares more leading bits with keys[first] than keys[last] does */
    const int common = __builtin_clzll(c->keys[first] ^ c->keys[last]);
    uint32_t lo = first, hi = last; /* invariant: keys[lo] shares more than `common` bits with keys[first] (trivially at lo = first), keys[hi] does not */
    while (hi - lo > 1)
    {
        const uint32_t mid = lo + (hi - lo) / 2; /* mid > first and keys are unique, so the xor is never 0 */
        if (__builtin_clzll(c->keys[first] ^ c->keys[mid]) > common) lo = mid; else hi = mid;
    }
    s->l = build(c, first, lo);
    s->r = build(c, lo + 1, last);
    s->lo.x = fminf(s->l->lo.x, s->r->lo.x); s->lo.y = fminf(s->l->lo.y, s->r->lo.y); s->lo.z = fminf(s->l->lo.z, s->r->lo.z); s->lo.w = 0.0f;
    s->hi.x = fmaxf(s->l->hi.x, s->r->hi.x); s->hi.y = fmaxf(s->l->hi.y, s->r->hi.y); s->hi.z = fmaxf(s->l->hi.z, s->r->hi.z); s->hi.w = 0.0f;
    const float area = half_area(s->lo, s->hi);
    const uint32_t count = last - first + 1;
    const float leafCost = area * (float)count;
    const float innerCost = (area * 2.0f + s->l->cost) + s->r->cost;
    const int collapse = count <= c->maxLeaf && leafCost <= innerCost;
    s->cost = collapse ? leafCost : innerCost;
    s->size = collapse ? 1u : 1u + s->l->size + s->r->size;
    s->leaf = collapse;
    return s;
}

static void emit(Ctx *c, const Sub *s, int32_t parent)
{
    const uint32_t ind = c->nOut++;
    Node *n = &c->out[ind];
    memset(n, 0, sizeof *n);
    n->bmin = s->lo; n->bmax = s->hi; n->bmin.w = n->bmax.w = 0.0f;
    n->parent = parent;
    if (s->leaf) { n->link = s->first; n->nPrims = (uint8_t)(s->last - s->first + 1); return; }
    emit(c, s->l, (int32_t)ind);
    n->link = c->nOut;
    emit(c, s->r, (int32_t)ind);
}

static void release(Sub *s) { if (!s) return; release(s->l); release(s->r); free(s); }

int port_build_lbvh(const Triangle *tris, uint32_t n, uint32_t maxLeaf, Node *nodes_out, uint32_t *n_nodes_out, uint32_t *indices_out)
{
    if (n == 0) return 1;
    f4 *pmin = (f4 *)malloc(sizeof(f4) * n), *pmax = (f4 *)malloc(sizeof(f4) * n);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * n);
    float lo[3] = {3.402823466e+38f, 3.402823466e+38f, 3.402823466e+38f}, hi[3] = {-3.402823466e+38f, -3.402823466e+38f, -3.402823466e+38f};
    for (uint32_t i = 0; i < n; i++)
    {
        const f4 a = tris[i].v0.p, b = tris[i].v1.p, c = tris[i].v2.p;
        pmin[i].x = fminf(fminf(a.x, b.x), c.x); pmin[i].y = fminf(fminf(a.y, b.y), c.y); pmin[i].z = fminf(fminf(a.z, b.z), c.z); pmin[i].w = 0.0f;
        pmax[i].x = fmaxf(fmaxf(a.x, b.x), c.x); pmax[i].y = fmaxf(fmaxf(a.y, b.y), c.y); pmax[i].z = fmaxf(fmaxf(a.z, b.z), c.z); pmax[i].w = 0.0f;
        const float cx = (pmin[i].x + pmax[i].x) * 0.5f, cy = (pmin[i].y + pmax[i].y) * 0.5f, cz = (pmin[i].z + pmax[i].z) * 0.5f;
        lo[0] = fminf(lo[0], cx); lo[1] = fminf(lo[1], cy); lo[2] = fminf(lo[2], cz);
        hi[0] = fmaxf(hi[0], cx); hi[1] = fmaxf(hi[1], cy); hi[2] = fmaxf(hi[2], cz);
    }
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t x = quantize10((pmin[i].x + pmax[i].x) * 0.5f, lo[0], hi[0]), y = quantize10((pmin[i].y + pmax[i].y) * 0.5f, lo[1], hi[1]),
                       z = quantize10((pmin[i].z + pmax[i].z) * 0.5f, lo[2], hi[2]);
        const uint32_t m = (spread10(x) << 2) | (spread10(y) << 1) | spread10(z);
        keys[i] = ((uint64_t)m << 32) | (uint64_t)i;
    }
    qsort(keys, n, sizeof(uint64_t), cmp_u64);
    for (uint32_t i = 0; i < n; i++) indices_out[i] = (uint32_t)(keys[i] & 0xffffffffu);
    Ctx c = {keys, pmin, pmax, maxLeaf, nodes_out, 0};
    Sub *root = build(&c, 0, n - 1);
    emit(&c, root, -1);
    *n_nodes_out = c.nOut;
    release(root);
    free(pmin); free(pmax); free(keys);
    return 0;
}
