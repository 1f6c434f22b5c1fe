"""Shared helpers of the parity tests: drive two contexts through the same loop and compare their state."""
import numpy as np

from fluctus_b200 import SLOT, Tracer

USED_SLOTS = [s for s in range(SLOT.COUNT) if s not in SLOT.UNUSED]


def slot_name(s):
    base = max(k for k in SLOT.NAMES if k <= s)
    return "%s[%d]" % (SLOT.NAMES[base], s - base)


def compare_tasks(a, b, what, n_live=None, exact=True):
    """a, b: (64, N) uint32 path-state dumps. Bit-exact on every slot the wavefront path writes, except lastPdfW of
    paths whose throughput is zero: the reference leaves pdfW unset there (src/glossy.cl:58-59 returns before writing it,
    src/wf_mat_*.cl:39 declares it uninitialised), so its value is whatever was on the stack."""
    n = a.shape[1] if n_live is None else n_live
    bad = []
    for s in USED_SLOTS:
        x, y = a[s, :n], b[s, :n]
        neq = x != y
        if s == SLOT.LAST_PDF_W:
            t_zero = (a[SLOT.T, :n].view(np.float32) == 0) & (a[SLOT.T + 1, :n].view(np.float32) == 0) & (a[SLOT.T + 2, :n].view(np.float32) == 0)
            neq &= ~t_zero
        # NaN payloads: compare as floats too (NaN == NaN for our purposes when both are NaN)
        if neq.any():
            xf, yf = x.view(np.float32), y.view(np.float32)
            both_nan = np.isnan(xf) & np.isnan(yf)
            neq &= ~both_nan
        if neq.any():
            idx = np.flatnonzero(neq)
            bad.append("%s: %d/%d differ, first path %d: %r vs %r (0x%08x vs 0x%08x)" % (
                slot_name(s), len(idx), n, idx[0], x.view(np.float32)[idx[0]], y.view(np.float32)[idx[0]], x[idx[0]], y[idx[0]]))
    assert not bad, "%s: path state differs\n  " % what + "\n  ".join(bad)


def compare_counters(ca, cb, what):
    da, db = ca.as_dict(), cb.as_dict()
    assert da == db, "%s: queue counters differ: %r vs %r" % (what, da, db)


def compare_queues(ctx_a, ctx_b, cnt, what):
    d = cnt.as_dict()
    for q in ("raygen", "extension", "shadow", "diffuse", "glossy", "ggxRefl", "ggxRefr", "delta"):
        n = d[q + "Queue"]
        qa, qb = ctx_a.readQueue(q, n), ctx_b.readQueue(q, n)
        if q == "raygen":  # order decides the pixel each regenerated path gets (wf_raygen.cl:25)
            assert np.array_equal(qa, qb), "%s: raygen queue order differs" % what
        else:
            assert np.array_equal(np.sort(qa), np.sort(qb)), "%s: %s queue holds different paths" % (what, q)


def compare_pixels(pa, pb, what, rtol=1e-4, exact_rgb=False):
    """pa, pb: (P, 4) accumulators. Alpha (sample count) must match exactly; RGB within rtol of the oracle
    (|a-b| <= rtol * max(|b|, 1e-3)) -- bit-exact when at most one path per pixel terminates per iteration."""
    assert np.array_equal(pa[:, 3], pb[:, 3]), "%s: per-pixel sample counts differ (%d pixels)" % (what, int((pa[:, 3] != pb[:, 3]).sum()))
    if exact_rgb:
        neq = (pa[:, :3].view(np.uint32) != pb[:, :3].view(np.uint32)) & ~(np.isnan(pa[:, :3]) & np.isnan(pb[:, :3]))
        assert not neq.any(), "%s: %d pixel channels not bit-identical" % (what, int(neq.sum()))
        return 0.0
    err = np.abs(pa[:, :3].astype(np.float64) - pb[:, :3]) / np.maximum(np.abs(pb[:, :3]).astype(np.float64), 1e-3)
    worst = float(np.nanmax(err)) if err.size else 0.0
    assert worst <= rtol, "%s: max relative radiance error %.3g > %.1g (%d channels over)" % (what, worst, rtol, int((err > rtol).sum()))
    return worst


def setup_context(ctx, scene, params, env=None):
    ctx.uploadSceneData(scene)
    if env is not None:
        ctx.createEnvMap(env)
    ctx.setupPixelStorage(params.width, params.height)
    ctx.updateParams(params)
    return Tracer(ctx, params)


def run_lockstep(gpu, cpu, scene, params, iterations, env=None, check_every=1, exact_rgb=None):
    """Run the reference's loop on both contexts, comparing complete state after the prologue and after iterations."""
    tg, tc = setup_context(gpu, scene, params, env), setup_context(cpu, scene, params, env)
    tg.start()
    tc.start()
    compare_tasks(gpu.readTasks(), cpu.readTasks(), "after reset+raygen+extrays")
    # Two paths that hold the same pixel (the pixel counter wraps around the image while older paths are still in
    # flight, wf_raygen.cl:25) may terminate in the same iteration; the order of their float atomics is free, so RGB is
    # compared to 1e-5 relative (far inside the 1e-4 of the north star) unless the caller knows better.
    if exact_rgb is None:
        exact_rgb = False
    for it in range(iterations):
        cg, cc = tg.iterate(), tc.iterate()
        what = "iteration %d" % it
        compare_counters(cg, cc, what)
        if it % check_every == 0 or it == iterations - 1:
            compare_queues(gpu, cpu, cc, what)
            compare_tasks(gpu.readTasks(), cpu.readTasks(), what)
            pg, pc = gpu.readPixels(), cpu.readPixels()
            compare_pixels(pg, pc, what, rtol=1e-5, exact_rgb=exact_rgb)
            if hasattr(gpu, "readPreview") and hasattr(cpu, "readPreview"):  # display pass (mk_postprocess.cl), run by iterate()
                vg, vc = gpu.readPreview(), cpu.readPreview()
                same_in = np.array_equal(pg.view(np.uint32), pc.view(np.uint32))
                compare_pixels(vg, vc, what + " (post-processed preview)", rtol=1e-5, exact_rgb=same_in)
    return tg, tc


# ---------------------------------------------------------------------------------------------- microkernel integrator
def compare_mk_tasks(a, b, what, n_live):
    """a (ours), b (reference kernels): (64, N) path-state dumps of the microkernel integrator; bit-exact on every slot,
    phase word included, for the first n_live = min(W*H, N) paths.  Two masks, both for values the reference leaves
    UNINITIALISED: when a BSDF sampler rejects its direction without writing pdfW (src/glossy.cl:58-59; declared without
    initialiser at src/mk_sample_bsdf.cl:160) the path terminates, but the garbage still lands in lastPdfW and, through
    T * bsdf * costh / pdfW with bsdf = 0, in T (0, -0 or NaN).  This repo pins pdfW = 0 there, so lastPdfW is skipped where
    ours is exactly 0 and T where ours is all zero/NaN.  Nothing reads either before it is overwritten."""
    n = n_live
    bad = []
    ta = a[SLOT.T:SLOT.T + 3, :n].view(np.float32)
    t_tainted = (np.isnan(ta) | (ta == 0)).all(axis=0)
    for s in USED_SLOTS + [SLOT.PHASE]:
        x, y = a[s, :n], b[s, :n]
        neq = x != y
        if s == SLOT.LAST_PDF_W:
            neq &= ~(x.view(np.float32) == 0)
        if SLOT.T <= s < SLOT.T + 3:
            neq &= ~t_tainted
        if neq.any():
            xf, yf = x.view(np.float32), y.view(np.float32)
            neq &= ~(np.isnan(xf) & np.isnan(yf))
        if neq.any():
            idx = np.flatnonzero(neq)
            nm = "phase" if s == SLOT.PHASE else slot_name(s)
            bad.append("%s: %d/%d differ, first path %d: %r vs %r (0x%08x vs 0x%08x)" % (
                nm, len(idx), n, idx[0], x.view(np.float32)[idx[0]], y.view(np.float32)[idx[0]], x[idx[0]], y[idx[0]]))
    assert not bad, "%s: microkernel path state differs\n  " % what + "\n  ".join(bad)


def mk_stats(ctx):
    s = ctx.getStats()
    return (int(s.primaryRays) & 0xffffffff, int(s.extensionRays) & 0xffffffff, int(s.shadowRays) & 0xffffffff, int(s.samples) & 0xffffffff)


def run_mk_lockstep(gpu, cpu, scene, params, spp, env=None, check_every=1, interactive=False):
    """Drive two contexts through the reference's microkernel loop -- Tracer::renderSingle (src/tracer.cpp:112-150), or with
    interactive=True the preview + progressive calls of Tracer::update (src/tracer.cpp:267-299) -- kernel by kernel,
    comparing the complete path state after every enqueue and the accumulator after every splat."""
    tg, tc = setup_context(gpu, scene, params, env), setup_context(cpu, scene, params, env)
    n_live = min(params.width * params.height, gpu.getNumTasks())
    for c in (gpu, cpu):
        c.resetStats()
    step = [0]

    def both(method, what):
        for c in (gpu, cpu):
            getattr(c, method)(params)
            c.finishQueue()
        step[0] += 1
        if step[0] % check_every == 0 or method.startswith("enqueueSplat"):
            compare_mk_tasks(gpu.readTasks(), cpu.readTasks(), what, n_live)

    def splat(method, what):
        both(method, what)
        pg, pc = gpu.readPixels(), cpu.readPixels()
        compare_pixels(pg, pc, what, exact_rgb=True)  # path g owns pixel g: no atomics, so bit-exact
        both("enqueuePostprocessKernel", what + " display pass")
        compare_pixels(gpu.readPreview(), cpu.readPreview(), what + " (post-processed preview)", exact_rgb=True)

    both("enqueueResetKernel", "after mk reset")
    if interactive:
        both("enqueueRayGenKernel", "preview raygen")
        for seg in range(2):
            both("enqueueNextVertexKernel", "preview nextVertex %d" % seg)
            both("enqueueBsdfSampleKernel", "preview sampleBsdf %d" % seg)
        splat("enqueueSplatPreviewKernel", "preview splat")
        for it in range(spp):
            both("enqueueRayGenKernel", "update %d raygen" % it)
            both("enqueueNextVertexKernel", "update %d nextVertex" % it)
            both("enqueueBsdfSampleKernel", "update %d sampleBsdf" % it)
            splat("enqueueSplatKernel", "update %d splat" % it)
    else:
        for s in range(spp):
            both("enqueueRayGenKernel", "sample %d raygen" % s)
            for bounce in range(params.maxBounces + 1):
                both("enqueueNextVertexKernel", "sample %d bounce %d nextVertex" % (s, bounce))
                both("enqueueBsdfSampleKernel", "sample %d bounce %d sampleBsdf" % (s, bounce))
            splat("enqueueSplatKernel", "sample %d splat" % s)
    assert mk_stats(gpu) == mk_stats(cpu), "ray/sample statistics differ: %r vs %r" % (mk_stats(gpu), mk_stats(cpu))
    return tg, tc


# ---------------------------------------------------------------------------------------------- hierarchy checks
def validate_bvh(nodes, indices, tris, max_leaf=8, unique_refs=True):
    """Structural contract of the reference's flattened hierarchy (src/bvhnode.hpp:50-59, src/sbvh.cpp:52-73) that the traversal
    relies on: depth-first order with left child = self + 1, rightChild inside the array, parent links, every node box
    containing its children / triangles, leaves of 1..max_leaf references, and (for builders without spatial splits) every
    triangle referenced exactly once.  Returns (depth, number of leaves, SAH cost normalised by the root area)."""
    n = len(nodes)
    link, nprims, parent = nodes["link"].astype(np.int64), nodes["nPrims"].astype(np.int64), nodes["parent"].astype(np.int64)
    bmin, bmax = nodes["bmin"][:, :3], nodes["bmax"][:, :3]
    inner = np.flatnonzero(nprims == 0)
    leaves = np.flatnonzero(nprims > 0)
    assert len(inner) + 1 == len(leaves), "a full binary tree has one more leaf than inner nodes"
    left, right = inner + 1, link[inner]
    assert (right > left).all() and (right < n).all(), "rightChild out of range"
    assert parent[0] == -1 and (parent[left] == inner).all() and (parent[right] == inner).all(), "parent links"
    for ch in (left, right):
        assert (bmin[ch] >= bmin[inner]).all() and (bmax[ch] <= bmax[inner]).all(), "child box not inside its parent's"
    assert (nprims[leaves] <= max_leaf).all()
    assert (link[leaves] + nprims[leaves] <= len(indices)).all()
    starts = link[leaves]
    order = np.argsort(starts)
    assert (starts[order][1:] == (starts[order] + nprims[leaves][order])[:-1]).all() and starts[order][0] == 0, "leaf ranges must tile the index list"
    assert starts[order][-1] + nprims[leaves][order][-1] == len(indices)
    if unique_refs:
        assert len(indices) == len(tris) and np.array_equal(np.sort(indices), np.arange(len(tris), dtype=indices.dtype)), "every triangle exactly once"
    # triangles inside their leaf's box (not for spatial-split builders: an SBVH reference is clipped to its leaf, src/sbvh.cpp:268-330)
    if unique_refs:
        leaf_of_ref = np.repeat(leaves[order], nprims[leaves][order])
        P = np.stack([tris["v0"]["p"][:, :3], tris["v1"]["p"][:, :3], tris["v2"]["p"][:, :3]], axis=1)[indices]  # (refs, 3 vertices, xyz)
        assert (P.min(axis=1) >= bmin[leaf_of_ref]).all() and (P.max(axis=1) <= bmax[leaf_of_ref]).all(), "triangle outside its leaf box"
    # depth (iteratively, parents precede children in DFS order) and SAH cost with the reference's constants
    depth = np.zeros(n, np.int64)
    for i in range(1, n):
        depth[i] = depth[parent[i]] + 1
    d = (bmax - bmin).astype(np.float64)
    area = d[:, 0] * d[:, 1] + d[:, 1] * d[:, 2] + d[:, 2] * d[:, 0]
    sah = (2.0 * area[inner].sum() + (area[leaves] * nprims[leaves]).sum()) / max(area[0], 1e-300)
    return int(depth.max()), len(leaves), float(sah)
