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

extern float g_ploc_tri_cost;
static Sub *build(const Ctx *c, uint32_t first, uint32_t last)
{
    Sub *s = (Sub *)calloc(1, sizeof(Sub));
    s->first = first; s->last = last;
    if (first == last)
    {
        const uint32_t tri = (uint32_t)(c->keys[first] & 0xffffffffu);
        s->lo = c->pmin[tri]; s->hi = c->pmax[tri]; s->cost = half_area(s->lo, s->hi) * g_ploc_tri_cost; s->size = 1; s->leaf = 1;
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
    const float leafCost = (area * (float)count) * g_ploc_tri_cost;
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
float g_ploc_tri_cost = 1.0f; /* SAH cost of one triangle test relative to one box test in the collapse decision (FLX_TUNE_BVH_TRI_COST / 100) */
void port_set_tri_cost(float c) { g_ploc_tri_cost = c; }
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

/* ---------------------------------------------------------------- parallel reinsertion (FLX_BVH_PLOC_OPT)
 * Post-pass over the finished PLOC tree, after Meister & Bittner, "Parallel reinsertion for bounding volume hierarchy optimization"
 * (2018): what top-down builders get from spatial splits -- little overlap between siblings -- a bottom-up builder can approach by
 * moving subtrees to where they enlarge the fewest boxes.  Per iteration, on the tree as it stands (restated here in the order the
 * GPU kernels of flx_bvh_build.cuh run, so that the result is the same bit for bit):
 *   search  every node x (not the root, not a child of the root, not inside a collapsed leaf) looks for the node y next to which it
 *           would sit best: x and its parent p leave (the sibling s takes p's place), p comes back as the parent of (y, x).  The gain
 *           is the half-area that inner nodes lose (p itself, and the ancestors of p below the lowest common ancestor "pivot", which
 *           shrink) minus what they gain (the new box of p, and the ancestors of y below the pivot, which grow).  The pivot walks from
 *           p to the root; under each pivot the subtree on the other side is searched depth first, left child first, pruned where even
 *           a perfect fit (direct cost = area of x) cannot beat the best gain so far;
 *   lock    a move with gain > 0 puts (gain bits << 32 | x) on the six nodes whose links it rewrites -- x, p, s, the grandparent g, y and
 *           y's parent -- with max(): the largest gain wins a contested node; a move that holds all six is a candidate;
 *   guard   simultaneous moves with disjoint link sets could still close a cycle (x1 goes below x2 while x2 goes below x1, or longer
 *           chains): that takes, for every move of the chain, another moving node strictly between its pivot and its target -- so a
 *           candidate with another candidate's x on its way down from the pivot to y stands back;
 *   apply   the remaining moves are carried out (disjoint link sets: no two write the same node);
 *   refit   boxes bottom-up.
 * After the last iteration cost / size / triangle count / collapse decision are recomputed bottom-up with the builder's formulas. */
static int g_ploc_reinsert = 0, g_ploc_depth_limit = 62;
void port_set_reinsert(int iterations) { g_ploc_reinsert = iterations; }
void port_set_depth_limit(int levels) { g_ploc_depth_limit = levels; }
#define RI_STACK 128
static void box_union(f4 alo, f4 ahi, f4 blo, f4 bhi, f4 *lo, f4 *hi)
{
    lo->x = fminf(alo.x, blo.x); lo->y = fminf(alo.y, blo.y); lo->z = fminf(alo.z, blo.z); lo->w = 0.0f;
    hi->x = fmaxf(ahi.x, bhi.x); hi->y = fmaxf(ahi.y, bhi.y); hi->z = fmaxf(ahi.z, bhi.z); hi->w = 0.0f;
}
static int ri_is_leaf(const PNode *nd, uint32_t n, int id) { return (uint32_t)id < n || nd[id].collapsed; }

static void ri_search(const PNode *nd, const int *parent, uint32_t n, int root, int x, float *gainOut, int *outOut, int *pivotOut)
{
    const f4 xlo = nd[x].lo, xhi = nd[x].hi;
    const float aX = half_area(xlo, xhi);
    const int p = parent[x];
    float best = 0.0f; int bestOut = -1, bestPivot = -1;
    float dDec = half_area(nd[p].lo, nd[p].hi);
    f4 pathLo = xlo, pathHi = xhi; /* set when the pivot leaves p */
    int child = x, pivot = p;
    int stackNode[RI_STACK]; float stackInc[RI_STACK];
    for (;;)
    {
        const int other = nd[pivot].left == child ? nd[pivot].right : nd[pivot].left;
        int sp = 0;
        stackNode[sp] = other; stackInc[sp] = 0.0f; sp++;
        while (sp > 0)
        {
            --sp;
            const int y = stackNode[sp]; const float inc = stackInc[sp];
            f4 ulo, uhi;
            box_union(nd[y].lo, nd[y].hi, xlo, xhi, &ulo, &uhi);
            const float direct = half_area(ulo, uhi);
            const float gain = (dDec - inc) - direct;
            if (gain > best) { best = gain; bestOut = y; bestPivot = pivot; }
            if (!ri_is_leaf(nd, n, y))
            {
                const float incChild = (inc + direct) - half_area(nd[y].lo, nd[y].hi);
                if (((dDec - incChild) - aX) > best && sp + 2 <= RI_STACK)
                {
                    stackNode[sp] = nd[y].right; stackInc[sp] = incChild; sp++;
                    stackNode[sp] = nd[y].left; stackInc[sp] = incChild; sp++;
                }
            }
        }
        if (pivot == root) break;
        if (pivot == p) { pathLo = nd[other].lo; pathHi = nd[other].hi; }
        else
        {
            box_union(pathLo, pathHi, nd[other].lo, nd[other].hi, &pathLo, &pathHi);
            dDec = dDec + (half_area(nd[pivot].lo, nd[pivot].hi) - half_area(pathLo, pathHi));
        }
        child = pivot; pivot = parent[pivot];
    }
    *gainOut = best; *outOut = bestOut; *pivotOut = bestPivot;
}

/* visits the nodes whose links a move rewrites: x, its parent p, its sibling s, its grandparent g, the target y and y's parent */
#define RI_FOR_LINKSET(BODY)                                                                                  \
    do {                                                                                                      \
        const int p_ = parent[x]; const int s_ = nd[p_].left == x ? nd[p_].right : nd[p_].left;               \
        int v; v = x; BODY; v = p_; BODY; v = s_; BODY; v = parent[p_]; BODY; v = y; BODY; v = parent[y]; BODY; \
    } while (0)

static void ri_refit_boxes(PNode *nd, uint32_t n, int id)
{
    if (ri_is_leaf(nd, n, id)) return;
    ri_refit_boxes(nd, n, nd[id].left); ri_refit_boxes(nd, n, nd[id].right);
    box_union(nd[nd[id].left].lo, nd[nd[id].left].hi, nd[nd[id].right].lo, nd[nd[id].right].hi, &nd[id].lo, &nd[id].hi);
}
static void ri_recost(PNode *nd, uint32_t n, uint32_t maxLeaf, int id)
{
    if (ri_is_leaf(nd, n, id)) return;
    const int l = nd[id].left, r = nd[id].right;
    ri_recost(nd, n, maxLeaf, l); ri_recost(nd, n, maxLeaf, r);
    PNode *o = &nd[id];
    box_union(nd[l].lo, nd[l].hi, nd[r].lo, nd[r].hi, &o->lo, &o->hi);
    const float area = half_area(o->lo, o->hi);
    const uint32_t count = nd[l].prims + nd[r].prims;
    const float leafCost = (area * (float)count) * g_ploc_tri_cost;
    const float innerCost = (area * 2.0f + nd[l].cost) + nd[r].cost;
    const int collapse = count <= maxLeaf && leafCost <= innerCost;
    o->cost = collapse ? leafCost : innerCost; o->size = collapse ? 1u : 1u + nd[l].size + nd[r].size; o->prims = count; o->collapsed = collapse;
}
static void ri_mark_alive(const PNode *nd, uint32_t n, int id, uint8_t *alive)
{
    alive[id] = 1;
    if (ri_is_leaf(nd, n, id)) return;
    ri_mark_alive(nd, n, nd[id].left, alive); ri_mark_alive(nd, n, nd[id].right, alive);
}
uint32_t g_ploc_reinsert_moves[64]; /* moves carried out per iteration of the last build (diagnostics) */
static void ploc_reinsert(PNode *nd, uint32_t n, uint32_t total, int root, int iterations)
{
    int *parent = (int *)malloc(sizeof(int) * total);
    uint8_t *alive = (uint8_t *)calloc(total, 1);
    float *gain = (float *)malloc(sizeof(float) * total);
    int *out = (int *)malloc(sizeof(int) * total), *piv = (int *)malloc(sizeof(int) * total);
    uint64_t *lock = (uint64_t *)malloc(sizeof(uint64_t) * total);
    for (uint32_t i = 0; i < total; i++) parent[i] = -1;
    ri_mark_alive(nd, n, root, alive);
    for (uint32_t i = 0; i < total; i++) if (alive[i] && !ri_is_leaf(nd, n, (int)i)) { parent[nd[i].left] = (int)i; parent[nd[i].right] = (int)i; }
    for (int it = 0; it < iterations; it++)
    {
        for (uint32_t i = 0; i < total; i++) { gain[i] = 0.0f; out[i] = -1; lock[i] = 0; }
        for (uint32_t i = 0; i < total; i++)
            if (alive[i] && (int)i != root && parent[i] != root)
                ri_search(nd, parent, n, root, (int)i, &gain[i], &out[i], &piv[i]);
        for (uint32_t i = 0; i < total; i++)
            if (out[i] >= 0)
            {
                const int x = (int)i, y = out[i];
                uint32_t bits; memcpy(&bits, &gain[i], 4);
                const uint64_t key = ((uint64_t)bits << 32) | (uint64_t)i;
                RI_FOR_LINKSET(if (lock[v] < key) lock[v] = key);
            }
        uint32_t moves = 0;
        uint8_t *cand = (uint8_t *)calloc(total, 1), *win = (uint8_t *)calloc(total, 1);
        for (uint32_t i = 0; i < total; i++)
            if (out[i] >= 0)
            {
                const int x = (int)i, y = out[i];
                uint32_t bits; memcpy(&bits, &gain[i], 4);
                const uint64_t key = ((uint64_t)bits << 32) | (uint64_t)i;
                int ok = 1;
                RI_FOR_LINKSET(if (lock[v] != key) ok = 0);
                cand[i] = (uint8_t)ok;
            }
        for (uint32_t i = 0; i < total; i++)
            if (cand[i])
            {
                int ok = 1;
                for (int v = parent[out[i]]; v != piv[i]; v = parent[v])
                    if (cand[v]) ok = 0;
                win[i] = (uint8_t)ok;
            }
        free(cand);
        for (uint32_t i = 0; i < total; i++)
            if (win[i])
            {
                const int x = (int)i, y = out[i];
                const int p = parent[x], s = nd[p].left == x ? nd[p].right : nd[p].left, g = parent[p];
                if (nd[g].left == p) nd[g].left = s; else nd[g].right = s;
                parent[s] = g;
                const int yp = parent[y];
                if (nd[yp].left == y) nd[yp].left = p; else nd[yp].right = p;
                parent[p] = yp;
                nd[p].left = y; nd[p].right = x;
                parent[y] = p; parent[x] = p;
                moves++;
            }
        free(win);
        if (it < 64) g_ploc_reinsert_moves[it] = moves;
        ri_refit_boxes(nd, n, root);
        if (moves == 0) break;
    }
    free(parent); free(alive); free(gain); free(out); free(piv); free(lock);
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
        nd[j].cost = half_area(nd[j].lo, nd[j].hi) * g_ploc_tri_cost; nd[j].size = 1; nd[j].prims = 1; nd[j].collapsed = 0;
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
            const float leafCost = (area * (float)count) * g_ploc_tri_cost;
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
    const int reinserted = g_ploc_reinsert > 0 && n > 2;
    if (reinserted)
    {
        ploc_reinsert(nd, n, nextId, (int)cid[0], g_ploc_reinsert);
        ri_recost(nd, n, maxLeaf, (int)cid[0]);
    }
    uint32_t nOut = 0, nIdx = 0;
    ploc_emit(nd, cid[0], -1, nodes_out, &nOut, indices_out, &nIdx, keys, n, 0);
    *n_nodes_out = nOut;
    free(pmin); free(pmax); free(keys); free(nd); free(cid); free(next); free(nn);
    if (nIdx != n) return 3;
    /* depth of the emitted tree (parents precede children); beyond the limit an optimised tree falls back to the one it started from,
     * as flx_build_bvh does, and a plain one is an error there (5 here) */
    uint32_t *depth = (uint32_t *)calloc(nOut, sizeof(uint32_t)), deepest = 0;
    for (uint32_t i = 1; i < nOut; i++) { depth[i] = depth[nodes_out[i].parent] + 1u; if (depth[i] > deepest) deepest = depth[i]; }
    free(depth);
    if (deepest > (uint32_t)g_ploc_depth_limit)
    {
        if (!reinserted) return 5;
        const int keep = g_ploc_reinsert;
        g_ploc_reinsert = 0;
        const int rc = port_build_ploc(tris, n, maxLeaf, nodes_out, n_nodes_out, indices_out);
        g_ploc_reinsert = keep;
        return rc;
    }
    return 0;
}


/* ================================================================ reference pre-splitting (FLX_BVH_PLOC_SPLIT)
 * "Early split clipping" (Ernst & Greiner 2007) in front of the PLOC builder: a triangle whose box is large is handed to the builder
 * as several REFERENCES, each with the box of the part of the triangle inside one cell of a recursive spatial-median subdivision of
 * its box -- what the reference's SBVH gets from spatial splits (src/sbvh.cpp:118-142, 240-330): a long or large triangle no longer
 * forces a large box on every node above it.  The index list then names a triangle once per reference (n_indices > n_tris), as the
 * reference's SBVH does.
 *   box of a part = union of the triangle's vertices on that side of the plane and the points where its edges cross the plane,
 *   intersected with the cell (the SBVH's splitReference does the same, src/sbvh.cpp:268-330); split axis = longest extent of the
 *   cell, plane = its middle; a cell is split while its half-area exceeds `threshold` and depth < ESC_MAX_DEPTH.
 * References of one triangle are emitted depth first, lower side first. */
#define ESC_MAX_DEPTH 6
typedef struct { f4 lo, hi; int depth; } Cell;

static float axis_of(f4 v, int a) { return a == 0 ? v.x : (a == 1 ? v.y : v.z); }
static void set_axis(f4 *v, int a, float x) { if (a == 0) v->x = x; else if (a == 1) v->y = x; else v->z = x; }
static void grow(f4 *lo, f4 *hi, f4 p)
{
    lo->x = fminf(lo->x, p.x); lo->y = fminf(lo->y, p.y); lo->z = fminf(lo->z, p.z);
    hi->x = fmaxf(hi->x, p.x); hi->y = fmaxf(hi->y, p.y); hi->z = fmaxf(hi->z, p.z);
}

/* splits the part of triangle (a, b, c) inside `cell` at plane axis = pos; returns a bit mask: 1 = lower part exists, 2 = upper */
static int split_cell(f4 a, f4 b, f4 c, const Cell *cell, int axis, float pos, Cell *lower, Cell *upper)
{
    const float BIG = 3.402823466e+38f;
    f4 llo = {BIG, BIG, BIG, 0.0f}, lhi = {-BIG, -BIG, -BIG, 0.0f}, ulo = llo, uhi = lhi;
    const f4 v[3] = {a, b, c};
    for (int e = 0; e < 3; e++)
    {
        const f4 v0 = v[e], v1 = v[(e + 1) % 3];
        const float p0 = axis_of(v0, axis), p1 = axis_of(v1, axis);
        if (p0 <= pos) grow(&llo, &lhi, v0);
        if (p0 >= pos) grow(&ulo, &uhi, v0);
        if ((p0 < pos && p1 > pos) || (p0 > pos && p1 < pos))
        {
            const float t = (pos - p0) / (p1 - p0);
            f4 x;
            x.x = v0.x + t * (v1.x - v0.x); x.y = v0.y + t * (v1.y - v0.y); x.z = v0.z + t * (v1.z - v0.z); x.w = 0.0f;
            set_axis(&x, axis, pos);
            grow(&llo, &lhi, x);
            grow(&ulo, &uhi, x);
        }
    }
    set_axis(&lhi, axis, fminf(axis_of(lhi, axis), pos));
    set_axis(&ulo, axis, fmaxf(axis_of(ulo, axis), pos));
    /* intersect with the cell */
    llo.x = fmaxf(llo.x, cell->lo.x); llo.y = fmaxf(llo.y, cell->lo.y); llo.z = fmaxf(llo.z, cell->lo.z);
    lhi.x = fminf(lhi.x, cell->hi.x); lhi.y = fminf(lhi.y, cell->hi.y); lhi.z = fminf(lhi.z, cell->hi.z);
    ulo.x = fmaxf(ulo.x, cell->lo.x); ulo.y = fmaxf(ulo.y, cell->lo.y); ulo.z = fmaxf(ulo.z, cell->lo.z);
    uhi.x = fminf(uhi.x, cell->hi.x); uhi.y = fminf(uhi.y, cell->hi.y); uhi.z = fminf(uhi.z, cell->hi.z);
    int mask = 0;
    if (llo.x <= lhi.x && llo.y <= lhi.y && llo.z <= lhi.z) { lower->lo = llo; lower->hi = lhi; lower->depth = cell->depth + 1; mask |= 1; }
    if (ulo.x <= uhi.x && ulo.y <= uhi.y && ulo.z <= uhi.z) { upper->lo = ulo; upper->hi = uhi; upper->depth = cell->depth + 1; mask |= 2; }
    return mask;
}

/* the references of one triangle; out may be NULL (count only) */
static uint32_t esc_triangle(const Triangle *t, float threshold, f4 *rmin, f4 *rmax)
{
    Cell stack[2 * ESC_MAX_DEPTH + 2];
    int sp = 0;
    uint32_t count = 0;
    const f4 a = t->v0.p, b = t->v1.p, c = t->v2.p;
    Cell root;
    root.lo.x = fminf(fminf(a.x, b.x), c.x); root.lo.y = fminf(fminf(a.y, b.y), c.y); root.lo.z = fminf(fminf(a.z, b.z), c.z); root.lo.w = 0.0f;
    root.hi.x = fmaxf(fmaxf(a.x, b.x), c.x); root.hi.y = fmaxf(fmaxf(a.y, b.y), c.y); root.hi.z = fmaxf(fmaxf(a.z, b.z), c.z); root.hi.w = 0.0f;
    root.depth = 0;
    stack[sp++] = root;
    while (sp > 0)
    {
        const Cell cell = stack[--sp];
        int done = !(half_area(cell.lo, cell.hi) > threshold) || cell.depth >= ESC_MAX_DEPTH;
        if (!done)
        {
            const float ex = cell.hi.x - cell.lo.x, ey = cell.hi.y - cell.lo.y, ez = cell.hi.z - cell.lo.z;
            const int axis = (ex >= ey && ex >= ez) ? 0 : (ey >= ez ? 1 : 2);
            const float pos = (axis_of(cell.lo, axis) + axis_of(cell.hi, axis)) * 0.5f;
            Cell lower, upper;
            const int mask = split_cell(a, b, c, &cell, axis, pos, &lower, &upper);
            if (mask == 3)
            {
                stack[sp++] = upper; /* lower side first */
                stack[sp++] = lower;
                continue;
            }
            done = 1; /* the plane misses the part (rounding): keep the cell as it is */
        }
        if (rmin) { rmin[count] = cell.lo; rmax[count] = cell.hi; rmin[count].w = rmax[count].w = 0.0f; }
        count++;
    }
    return count;
}

/* threshold = alpha * half-area of the scene box (union of the triangle boxes).  Returns 0, or 4 when the references do not fit
 * indices_capacity (then *n_indices_out = the count needed). */
int port_build_ploc_split(const Triangle *tris, uint32_t n, uint32_t maxLeaf, float alpha, Node *nodes_out, uint32_t nodes_capacity, uint32_t *n_nodes_out,
                          uint32_t *indices_out, uint32_t indices_capacity, uint32_t *n_indices_out)
{
    if (n == 0) return 1;
    const float BIG = 3.402823466e+38f;
    f4 slo = {BIG, BIG, BIG, 0.0f}, shi = {-BIG, -BIG, -BIG, 0.0f};
    for (uint32_t i = 0; i < n; i++) { grow(&slo, &shi, tris[i].v0.p); grow(&slo, &shi, tris[i].v1.p); grow(&slo, &shi, tris[i].v2.p); }
    const float threshold = alpha * half_area(slo, shi);
    uint32_t *base = (uint32_t *)malloc(sizeof(uint32_t) * ((size_t)n + 1));
    base[0] = 0;
    for (uint32_t i = 0; i < n; i++) base[i + 1] = base[i] + esc_triangle(&tris[i], threshold, NULL, NULL);
    const uint32_t m = base[n];
    *n_indices_out = m;
    if (m > indices_capacity || 2u * m - 1u > nodes_capacity) { free(base); return 4; }
    f4 *rmin = (f4 *)malloc(sizeof(f4) * m), *rmax = (f4 *)malloc(sizeof(f4) * m);
    uint32_t *refTri = (uint32_t *)malloc(sizeof(uint32_t) * m);
    for (uint32_t i = 0; i < n; i++)
    {
        const uint32_t k = esc_triangle(&tris[i], threshold, rmin + base[i], rmax + base[i]);
        for (uint32_t j = 0; j < k; j++) refTri[base[i] + j] = i;
    }
    /* from here on the PLOC builder, over references instead of triangles: a pseudo triangle array would do, but the boxes are
     * what it reads, so build degenerate "triangles" whose box is the reference's */
    Triangle *pseudo = (Triangle *)calloc(m, sizeof(Triangle));
    for (uint32_t r = 0; r < m; r++) { pseudo[r].v0.p = rmin[r]; pseudo[r].v1.p = rmax[r]; pseudo[r].v2.p = rmin[r]; }
    uint32_t nNodes = 0;
    const int rc = port_build_ploc(pseudo, m, maxLeaf, nodes_out, &nNodes, indices_out);
    if (rc == 0) for (uint32_t k = 0; k < m; k++) indices_out[k] = refTri[indices_out[k]];
    *n_nodes_out = nNodes;
    free(pseudo); free(base); free(rmin); free(rmax); free(refTri);
    return rc;
}
