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

static void sorted_keys(const Triangle *tris, uint32_t n, f4 *pmin, f4 *pmax, uint64_t *keys)
{
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
}

int port_build_lbvh(const Triangle *tris, uint32_t n, uint32_t maxLeaf, Node *nodes_out, uint32_t *n_nodes_out, uint32_t *indices_out)
{
    if (n == 0) return 1;
    f4 *pmin = (f4 *)malloc(sizeof(f4) * n), *pmax = (f4 *)malloc(sizeof(f4) * n);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * n);
    sorted_keys(tris, n, pmin, pmax, keys);
    for (uint32_t i = 0; i < n; i++) indices_out[i] = (uint32_t)(keys[i] & 0xffffffffu);
    Ctx c = {keys, pmin, pmax, maxLeaf, nodes_out, 0};
    Sub *root = build(&c, 0, n - 1);
    emit(&c, root, -1);
    *n_nodes_out = c.nOut;
    release(root);
    free(pmin); free(pmax); free(keys);
    return 0;
}


/* ================================================================ PLOC (the FLX_BVH_PLOC builder of flx_bvh_build.cuh)
 * Locally-ordered clustering on the Morton order: every round each cluster picks, among the PLOC_RADIUS positions to either
 * side, the neighbour whose merged box has the smallest area (lowest position on ties); mutual pairs merge into a new node
 * that takes the lower partner's place; compaction keeps the order.  Node ids: leaf j (sorted position) -> j, inner nodes
 * n, n+1, ... in order of creation (by round, then by position).  Emission is the general depth-first one. */
#define PLOC_RADIUS 16
typedef struct { f4 lo, hi; int left, right; float cost; uint32_t size, prims; int collapsed; } PNode;

static void ploc_emit(const PNode *nd, uint32_t id, int32_t parent, Node *out, uint32_t *nOut, uint32_t *indices, uint32_t *nIdx, const uint64_t *keys, uint32_t n, int leafMode)
{
    /* leafMode: below a collapsed node only the triangles are listed, in order */
    if (leafMode)
    {
        if (id < n) indices[(*nIdx)++] = (uint32_t)(keys[id] & 0xffffffffu);
        else { ploc_emit(nd, (uint32_t)nd[id].left, 0, out, nOut, indices, nIdx, keys, n, 1); ploc_emit(nd, (uint32_t)nd[id].right, 0, out, nOut, indices, nIdx, keys, n, 1); }
        return;
    }
    const uint32_t ind = (*nOut)++;
    Node *o = &out[ind];
    memset(o, 0, sizeof *o);
    o->bmin = nd[id].lo; o->bmax = nd[id].hi; o->bmin.w = o->bmax.w = 0.0f;
    o->parent = parent;
    if (id < n || nd[id].collapsed)
    {
        o->link = *nIdx; o->nPrims = (uint8_t)nd[id].prims;
        ploc_emit(nd, id, 0, out, nOut, indices, nIdx, keys, n, 1);
        return;
    }
    ploc_emit(nd, (uint32_t)nd[id].left, (int32_t)ind, out, nOut, indices, nIdx, keys, n, 0);
    o->link = *nOut;
    ploc_emit(nd, (uint32_t)nd[id].right, (int32_t)ind, out, nOut, indices, nIdx, keys, n, 0);
}

int port_build_ploc(const Triangle *tris, uint32_t n, uint32_t maxLeaf, Node *nodes_out, uint32_t *n_nodes_out, uint32_t *indices_out)
{
    if (n == 0) return 1;
    f4 *pmin = (f4 *)malloc(sizeof(f4) * n), *pmax = (f4 *)malloc(sizeof(f4) * n);
    uint64_t *keys = (uint64_t *)malloc(sizeof(uint64_t) * n);
    sorted_keys(tris, n, pmin, pmax, keys);
    PNode *nd = (PNode *)calloc(2 * (size_t)n, sizeof(PNode));
    uint32_t *cid = (uint32_t *)malloc(sizeof(uint32_t) * n), *next = (uint32_t *)malloc(sizeof(uint32_t) * n);
    int *nn = (int *)malloc(sizeof(int) * n);
    for (uint32_t j = 0; j < n; j++)
    {
        const uint32_t tri = (uint32_t)(keys[j] & 0xffffffffu);
        nd[j].lo = pmin[tri]; nd[j].hi = pmax[tri]; nd[j].left = nd[j].right = -1;
        nd[j].cost = half_area(nd[j].lo, nd[j].hi) * 1.0f; nd[j].size = 1; nd[j].prims = 1; nd[j].collapsed = 0;
        cid[j] = j;
    }
    uint32_t m = n, nextId = n;
    while (m > 1)
    {
        for (int p = 0; p < (int)m; p++)
        {
            float best = 3.402823466e+38f; int bestq = -1;
            const PNode *a = &nd[cid[p]];
            for (int d = -PLOC_RADIUS; d <= PLOC_RADIUS; d++)
            {
                const int q = p + d;
                if (d == 0 || q < 0 || q >= (int)m) continue;
                const PNode *b = &nd[cid[q]];
                f4 lo, hi;
                lo.x = fminf(a->lo.x, b->lo.x); lo.y = fminf(a->lo.y, b->lo.y); lo.z = fminf(a->lo.z, b->lo.z); lo.w = 0.0f;
                hi.x = fmaxf(a->hi.x, b->hi.x); hi.y = fmaxf(a->hi.y, b->hi.y); hi.z = fmaxf(a->hi.z, b->hi.z); hi.w = 0.0f;
                const float area = half_area(lo, hi);
                if (area < best || bestq < 0) { best = area; bestq = q; }
            }
            nn[p] = bestq;
        }
        uint32_t kept = 0, merged = 0;
        for (int p = 0; p < (int)m; p++)
        {
            const int q = nn[p];
            const int mutual = q >= 0 && nn[q] == p;
            if (mutual && q < p) continue;                 /* the upper partner disappears */
            if (!(mutual && p < q)) { next[kept++] = cid[p]; continue; }
            const uint32_t id = nextId + merged++, l = cid[p], r = cid[q];
            PNode *o = &nd[id];
            o->lo.x = fminf(nd[l].lo.x, nd[r].lo.x); o->lo.y = fminf(nd[l].lo.y, nd[r].lo.y); o->lo.z = fminf(nd[l].lo.z, nd[r].lo.z); o->lo.w = 0.0f;
            o->hi.x = fmaxf(nd[l].hi.x, nd[r].hi.x); o->hi.y = fmaxf(nd[l].hi.y, nd[r].hi.y); o->hi.z = fmaxf(nd[l].hi.z, nd[r].hi.z); o->hi.w = 0.0f;
            const float area = half_area(o->lo, o->hi);
            const uint32_t count = nd[l].prims + nd[r].prims;
            const float leafCost = area * (float)count;
            const float innerCost = (area * 2.0f + nd[l].cost) + nd[r].cost;
            const int collapse = count <= maxLeaf && leafCost <= innerCost;
            o->left = (int)l; o->right = (int)r; o->cost = collapse ? leafCost : innerCost;
            o->size = collapse ? 1u : 1u + nd[l].size + nd[r].size; o->prims = count; o->collapsed = collapse;
            next[kept++] = id;
        }
        if (merged == 0) { free(pmin); free(pmax); free(keys); free(nd); free(cid); free(next); free(nn); return 2; }
        uint32_t *t = cid; cid = next; next = t;
        m = kept; nextId += merged;
    }
    uint32_t nOut = 0, nIdx = 0;
    ploc_emit(nd, cid[0], -1, nodes_out, &nOut, indices_out, &nIdx, keys, n, 0);
    *n_nodes_out = nOut;
    free(pmin); free(pmax); free(keys); free(nd); free(cid); free(next); free(nn);
    return nIdx == n ? 0 : 3;
}
