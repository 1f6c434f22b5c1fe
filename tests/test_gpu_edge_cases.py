"""GPU edge cases and size-independent properties: ragged sizes, degenerate scenes, the interactive first-iteration path,
API misuse, and -- at the BASELINE.json full size, where the serial oracle is too slow -- invariants of the loop and
bit-equality between the three traversal kernels."""
import ctypes as C
import os

import numpy as np
import pytest

from fluctus_b200 import CLContext, FluctusError, QueueCounters, SceneData, SLOT, Tracer, make_params, look_at
from fluctus_b200.scene import build_bvh, make_room_scene, room_params, _tri, _material
from fluctus_b200.structs import MATERIAL_DTYPE, TRIANGLE_DTYPE

from conftest import scene_blob
from parity_util import compare_counters, compare_pixels, compare_tasks, run_lockstep, setup_context

pytestmark = pytest.mark.gpu


def oracle_ctx(n):
    from oracle.oracle_host import PortContext, RefContext, port_available, ref_available
    if ref_available():
        return RefContext(n)
    if port_available():
        return PortContext(n)
    pytest.skip("no oracle library built")


@pytest.mark.parametrize("n_tasks", [1, 31, 33, 1237])
def test_ragged_task_counts(n_tasks):
    """NUM_TASKS that is not a multiple of the warp, the logic tile or the trace block; more pixels than paths."""
    scene = make_room_scene(materials="mixed")
    params = room_params(scene, 37, 23, max_bounces=3, separate_queues=True)
    with CLContext(n_tasks) as gpu:
        run_lockstep(gpu, oracle_ctx(n_tasks), scene, params, iterations=10)


def test_more_paths_than_pixels_and_one_pixel_image():
    scene = make_room_scene(materials="diffuse")
    for w, h, n in ((1, 1, 64), (5, 3, 600)):
        params = room_params(scene, w, h, max_bounces=2)
        with CLContext(n) as gpu:
            run_lockstep(gpu, oracle_ctx(n), scene, params, iterations=6)


def test_scene_whose_root_is_a_leaf():
    """Two triangles, one leaf, no inner node: the traversal starts at a leaf reference."""
    tris = np.array([_tri((-1, 0, -1), (1, 0, -1), (1, 0, 1), 0), _tri((-1, 0, -1), (1, 0, 1), (-1, 0, 1), 0)], TRIANGLE_DTYPE)
    nodes, indices = build_bvh(tris, max_leaf=4)
    assert len(nodes) == 1 and nodes[0]["nPrims"] == 2
    scene = SceneData(tris, indices, nodes, np.array([_material()], MATERIAL_DTYPE))
    cam = look_at((0.0, 1.5, 0.5), (0.0, 0.0, 0.0), fov=70.0)
    light = dict(pos=(0.0, 2.0, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(0.5, 0.5), E=(30.0, 30.0, 30.0))
    params = make_params(24, 16, cam, scene.world_radius, 2, light=light, max_bounces=3)
    with CLContext(24 * 16) as gpu:
        run_lockstep(gpu, oracle_ctx(24 * 16), scene, params, iterations=6)


@pytest.mark.parametrize("n_tasks", [700, 4000])
def test_interactive_first_iteration_path(n_tasks):
    """Tracer::update() with iteration == 0 (src/tracer.cpp:228-240): preview bounce clamp, three rounds with
    firstIteration = true (logic limited to min(W*H, N) paths, wf_logic.cl:45), pixel index advanced once."""
    scene = make_room_scene(materials="mixed")
    W, H = 48, 32  # 1536 pixels: one case with N < W*H, one with N > W*H
    params_g, params_c = room_params(scene, W, H, max_bounces=5), room_params(scene, W, H, max_bounces=5)
    cpu = oracle_ctx(n_tasks)
    with CLContext(n_tasks) as gpu:
        tg, tc = setup_context(gpu, scene, params_g), setup_context(cpu, scene, params_c)
        for it in range(5):
            cg, cc = tg.update(), tc.update()
            compare_counters(cg, cc, "update %d" % it)
            compare_tasks(gpu.readTasks(), cpu.readTasks(), "update %d" % it)
            compare_pixels(gpu.readPixels(), cpu.readPixels(), "update %d" % it, rtol=1e-5)
        assert tg.stats == tc.stats


def test_api_misuse_raises_like_the_reference():
    scene = make_room_scene(materials="diffuse")
    params = room_params(scene, 16, 16)
    with CLContext(256) as gpu:
        with pytest.raises(FluctusError, match="update_params|flx_update_params"):
            gpu.enqueueWfLogicKernel(params, False)  # nothing set up yet
        gpu.updateParams(params)
        with pytest.raises(FluctusError, match="resize"):
            gpu.enqueueWfResetKernel(params)
        gpu.setupPixelStorage(16, 16)
        with pytest.raises(FluctusError, match="upload_scene"):
            gpu.enqueueWfExtRayKernel(params)
        bad = SceneData(scene.tris, scene.indices, scene.nodes, scene.materials)
        bad.nodes["link"][0] = 10 ** 6  # child link out of range
        with pytest.raises(FluctusError, match="child links"):
            gpu.uploadSceneData(bad)
        bad = SceneData(scene.tris, scene.indices, scene.nodes, scene.materials)
        bad.tris["matId"][3] = 99
        with pytest.raises(FluctusError, match="material"):
            gpu.uploadSceneData(bad)
        bad = SceneData(scene.tris, scene.indices, scene.nodes, scene.materials)
        bad.indices[5] = 10 ** 6
        with pytest.raises(FluctusError, match="references triangle"):
            gpu.uploadSceneData(bad)
        gpu.uploadSceneData(scene)
        p2 = params.copy()
        p2.width = 32
        with pytest.raises(FluctusError, match="flx_resize"):
            gpu.updateParams(p2)
        # and it still works after all those errors
        tr = Tracer(gpu, params)
        tr.start()
        for _ in range(8):
            tr.iterate()
        assert gpu.readPixels()[:, 3].sum() > 0


def _full_size_conference():
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    return scene, conference_params(scene, 1920, 1080)


def test_full_size_invariants_and_kernel_variants_agree():
    """BASELINE metric configuration (conference 1920x1080, 8 bounces, N = 2^21) -- too big for the serial oracle, so:
    (1) the four traversal kernels (one ray per thread / persistent, one majority step per iteration [production] / round-1
        persistent + smem treelet / round-1 persistent) must leave bit-identical path state and counters,
    (2) every iteration the extension queue is a permutation of all paths and queue lengths are consistent,
    (3) every terminated path with at least one segment splats exactly once: sum of pixel weights == regenerated paths,
    (4) radiance is finite and non-negative."""
    scene, params = _full_size_conference()
    N, iters = 1 << 21, 12
    states, pixels, counters = [], [], []
    for variant in (0, 1, 2, 3):
        with CLContext(N) as gpu:
            gpu.setTuning(trace_variant=variant)
            tr = setup_context(gpu, scene, params)
            tr.start()
            regenerated = 0
            cs = []
            for it in range(iters):
                cnt = tr.iterate()
                cs.append(cnt.as_dict())
                regenerated += cnt.raygenQueue
                assert cnt.extensionQueue == N, "every live path is extended every iteration (SURVEY 3B)"
                assert cnt.diffuseQueue + cnt.raygenQueue == N and cnt.shadowQueue <= cnt.diffuseQueue
                if variant == 1 and it in (0, iters - 1):
                    q = gpu.readQueue("extension", N)
                    assert np.array_equal(np.sort(q), np.arange(N, dtype=np.uint32)), "extension queue is not a permutation"
                    rq = gpu.readQueue("raygen", cnt.raygenQueue)
                    assert (np.diff(rq.astype(np.int64)) > 0).all(), "raygen queue is not in ascending path order"
            pix = gpu.readPixels()
            # the paths regenerated in the LAST iteration have not terminated yet; all earlier ones that terminated have splatted
            assert np.isfinite(pix).all() and (pix >= 0).all()
            alive = gpu.readTasks()
            assert pix[:, 3].sum() == regenerated, "sum of sample weights must equal the number of terminated (= regenerated) paths"
            states.append(alive)
            pixels.append(pix)
            counters.append(cs)
    assert counters[0] == counters[1] == counters[2] == counters[3]
    for k in (1, 2, 3):
        compare_tasks(states[0], states[k], "trace variant 0 vs %d at full size" % k)
        compare_pixels(pixels[0], pixels[k], "trace variant 0 vs %d at full size" % k, rtol=1e-5)


def test_traversal_stack_placements_agree():
    """Where the traversal stack lives is invisible in the results: all of it in local memory (0, the default), its first 4 / 8 / 24
    levels in shared memory, or local memory with the newest entry kept in a register (-1) -- same path state bit for bit after 8
    iterations of the fused loop on Conference at 1280x720, N = 2^20."""
    from bench_configs import conference_params
    scene = SceneData.load_blob(scene_blob("conference"))
    params = conference_params(scene, 1280, 720)
    N = 1 << 20
    base = None
    for placement in (0, 4, 8, 24, -1):
        with CLContext(N) as gpu:
            gpu.setTuning(smem_stack=placement)
            tr = setup_context(gpu, scene, params)
            tr.start()
            tr.render(8)
            tasks, pix = gpu.readTasks(), gpu.readPixels()
        if base is None:
            base = (tasks, pix)
        else:
            compare_tasks(base[0], tasks, "stack placement 0 vs %d" % placement)
            compare_pixels(base[1], pix, "stack placement 0 vs %d" % placement, rtol=1e-5)


def test_full_size_is_deterministic_and_matches_fused_loop():
    """Two runs (one per-stage with host round trips, one fused flx_render with the shadow/extension overlap) at full size
    give the same path state bit for bit: no race decides anything that matters."""
    scene, params = _full_size_conference()
    N, iters = 1 << 21, 10
    with CLContext(N) as a, CLContext(N) as b:
        ta, tb = setup_context(a, scene, params), setup_context(b, scene, params)
        ta.start(); tb.start()
        for _ in range(iters):
            ta.iterate()
        b.resetStats()
        tb.render(iters)
        compare_tasks(a.readTasks(), b.readTasks(), "per-stage vs fused at full size")
        compare_pixels(a.readPixels(), b.readPixels(), "per-stage vs fused at full size", rtol=1e-5)
        st = b.getStats()
        assert (st.extensionRays, st.shadowRays, st.primaryRays) == (ta.stats["extensionRays"], ta.stats["shadowRays"], ta.stats["primaryRays"])


@pytest.mark.parametrize("separate", [False, True])
def test_every_stage_boundary_is_observable_despite_deferred_fusion(separate):
    """The per-stage ABI defers logic (+ raygen) so that logic, raygen, materials arriving back to back run as one fused kernel
    (flx_ctx::pendingStages).  An observer must not be able to tell: reading the path state after ANY single stage call has to
    show exactly what the reference's kernel sequence has produced up to that call.  Stage by stage against the oracle, with a
    read-back after every call (which forces the pending stages out as the separate kernels), interleaved with iterations that
    are left to fuse."""
    from fluctus_b200 import QueueCounters
    from parity_util import compare_counters, compare_tasks, setup_context
    from test_gpu_parity import oracle_ctx
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 64, 48, 4096
    params = room_params(scene, W, H, max_bounces=4, separate_queues=separate)
    with CLContext(N) as gpu:
        cpu = oracle_ctx(N)
        tg, tc = setup_context(gpu, scene, params), setup_context(cpu, scene, params)
        tg.start()
        tc.start()
        stages = ["enqueueWfLogicKernel", "enqueueWfRaygenKernel", "enqueueWfMaterialKernels", "enqueueWfExtRayKernel", "enqueueWfShadowRayKernel"]
        for it in range(6):
            if it % 2 == 1:  # an iteration nobody looks into: logic + raygen + materials fuse
                cg, cc = tg.iterate(), tc.iterate()
                compare_counters(cg, cc, "fused iteration %d" % it)
                compare_tasks(gpu.readTasks(), cpu.readTasks(), "after fused iteration %d" % it)
                continue
            for stage in stages:
                for c in (gpu, cpu):
                    if stage == "enqueueWfLogicKernel":
                        c.enqueueWfLogicKernel(params, False)
                    else:
                        getattr(c, stage)(params)
                if it % 4 == 2 and stage == "enqueueWfLogicKernel":
                    continue  # first look after raygen: logic AND raygen are pending and have to come out as two kernels
                compare_tasks(gpu.readTasks(), cpu.readTasks(), "iteration %d after %s" % (it, stage))
                compare_counters(gpu.readCounters(), cpu.readCounters(), "iteration %d after %s" % (it, stage))
            cnts = []
            for c, tr in ((gpu, tg), (cpu, tc)):
                cnt = c.readCounters()
                c.enqueueClearWfQueues()
                c.enqueuePostprocessKernel(params)
                c.finishQueue()
                c.updatePixelIndex(W * H, cnt.raygenQueue)
                cnts.append(cnt)
            compare_counters(cnts[0], cnts[1], "iteration %d" % it)
            from parity_util import compare_pixels
            compare_pixels(gpu.readPixels(), cpu.readPixels(), "iteration %d" % it, rtol=1e-5)


def test_checkpoint_resume_continues_bit_identically(tmp_path):
    """flx_checkpoint_save / _load: a render interrupted after 7 iterations and resumed in a NEW context ends with the same path
    state, queues, counters and statistics as the uninterrupted one (radiance up to the order of float atomics); a checkpoint
    of another shape is refused."""
    from parity_util import compare_pixels, compare_tasks, setup_context
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 64, 48, 4096
    params = room_params(scene, W, H, max_bounces=4, separate_queues=True)
    ck = tmp_path / "render.ckpt"
    with CLContext(N) as a:
        ta = setup_context(a, scene, params)
        ta.start()
        for _ in range(7):
            ta.iterate()
        a.saveCheckpoint(ck)
        a.render(3)
        for _ in range(4):
            ta.iterate()
        want_tasks, want_pix, want_stats = a.readTasks(), a.readPixels(), a.getStats()
        want_cnt = a.readCounters()
    with CLContext(N) as b:
        tb = setup_context(b, scene, params)
        b.loadCheckpoint(ck)
        b.render(3)
        for _ in range(4):
            tb.iterate()
        compare_tasks(b.readTasks(), want_tasks, "resumed render")
        compare_pixels(b.readPixels(), want_pix, "resumed render", rtol=1e-5)
        got = b.getStats()
        assert (got.extensionRays, got.shadowRays, got.primaryRays, got.iterations) == (want_stats.extensionRays, want_stats.shadowRays, want_stats.primaryRays, want_stats.iterations)
        assert b.readCounters().as_dict() == want_cnt.as_dict()
    with CLContext(N // 2) as c:
        setup_context(c, scene, params)
        with pytest.raises(FluctusError, match="checkpoint is for"):
            c.loadCheckpoint(ck)
        with pytest.raises(FluctusError):
            c.loadCheckpoint(tmp_path / "missing.ckpt")
    # a checkpoint continues the render it was written for: another scene, or another camera / light / sampling setup, is refused
    with CLContext(N) as d:
        other = make_room_scene(materials="diffuse", n_blobs=3)
        setup_context(d, other, room_params(other, W, H, max_bounces=4, separate_queues=True))
        with pytest.raises(FluctusError, match="different scene"):
            d.loadCheckpoint(ck)
        p2 = room_params(scene, W, H, max_bounces=5, separate_queues=True)
        setup_context(d, scene, p2)
        with pytest.raises(FluctusError, match="RenderParams"):
            d.loadCheckpoint(ck)
        p3 = room_params(scene, W, H, max_bounces=4, separate_queues=True)
        p3.ppParams.exposure = 2.5  # display-only: does not decide the accumulator, so the checkpoint still fits
        setup_context(d, scene, p3)
        d.loadCheckpoint(ck)
        # a damaged payload must not reach the device: a queue entry naming a path that does not exist, a counter beyond N
        raw = bytearray(open(ck, "rb").read())
        header = 8 + 8 * 4 + 16
        q0 = header + N * 64 * 4  # first entry of the raygen queue
        bad = bytearray(raw)
        bad[q0:q0 + 4] = (N + 7).to_bytes(4, "little")
        cnt_at = header + N * 64 * 4 + 8 * N * 4
        if int.from_bytes(raw[cnt_at:cnt_at + 4], "little") == 0:  # make sure that entry is live
            bad[cnt_at:cnt_at + 4] = (1).to_bytes(4, "little")
        (tmp_path / "bad_entry.ckpt").write_bytes(bad)
        with pytest.raises(FluctusError, match="names path"):
            d.loadCheckpoint(tmp_path / "bad_entry.ckpt")
        bad = bytearray(raw)
        bad[cnt_at + 4:cnt_at + 8] = (N + 1).to_bytes(4, "little")
        (tmp_path / "bad_counter.ckpt").write_bytes(bad)
        with pytest.raises(FluctusError, match="more than the"):
            d.loadCheckpoint(tmp_path / "bad_counter.ckpt")
        (tmp_path / "cut.ckpt").write_bytes(raw[:len(raw) // 2])
        with pytest.raises(FluctusError, match="truncated"):
            d.loadCheckpoint(tmp_path / "cut.ckpt")
        d.loadCheckpoint(ck)  # and the context still works
        d.render(2)
        assert np.isfinite(d.readPixels()).all()


def test_hierarchy_deeper_than_the_traversal_stack_is_refused(tmp_path):
    """The traversal keeps 64 stack entries (reference: uint stack[64], src/bvh.cl:240) and pushes without a bounds check, like the
    reference.  The in-repo builders stop at depth 62; a caller's Node[] -- or a cache file read by flx_hierarchy_import, which
    checks no structure -- can be deeper: a right-leaning chain of 80 inner nodes must be refused at upload, by the device repack
    and by the host repack alike, and a chain of 60 must still render."""
    from fluctus_b200.scene_io import export_hierarchy, import_hierarchy
    from fluctus_b200.structs import NODE_DTYPE

    def chain(depth):
        n_tris = depth + 1
        tris = np.array([_tri((k, 0, 0), (k + 0.9, 0, 0), (k, 0.9, 0), 0) for k in range(n_tris)], TRIANGLE_DTYPE)
        nodes = np.zeros(2 * depth + 1, NODE_DTYPE)
        # inner node 2k has the leaf 2k+1 on the left and the next inner node (or the last leaf) on the right
        for k in range(depth):
            i = 2 * k
            nodes[i]["bmin"][:3], nodes[i]["bmax"][:3] = (k, 0, 0), (n_tris, 0.9, 0)
            nodes[i]["parent"], nodes[i]["link"], nodes[i]["nPrims"] = (i - 2 if k else -1), i + 2, 0
            nodes[i + 1]["bmin"][:3], nodes[i + 1]["bmax"][:3] = (k, 0, 0), (k + 0.9, 0.9, 0)
            nodes[i + 1]["parent"], nodes[i + 1]["link"], nodes[i + 1]["nPrims"] = i, k, 1
        last = 2 * depth
        nodes[last]["bmin"][:3], nodes[last]["bmax"][:3] = (depth, 0, 0), (depth + 0.9, 0.9, 0)
        nodes[last]["parent"], nodes[last]["link"], nodes[last]["nPrims"] = last - 2, depth, 1
        return SceneData(tris, np.arange(n_tris, dtype=np.uint32), nodes, np.array([_material()], MATERIAL_DTYPE))

    deep, fine = chain(80), chain(60)
    cache = tmp_path / "deep.bin"
    export_hierarchy(cache, deep.nodes, deep.indices)
    nodes, indices = import_hierarchy(cache)  # the importer takes it: structure is the uploader's business
    assert len(nodes) == len(deep.nodes)
    with CLContext(256) as gpu:
        for on_host in (0, 1):
            gpu.setTuning(repack_on_host=on_host)
            with pytest.raises(FluctusError, match="levels deep"):
                gpu.uploadSceneData(SceneData(deep.tris, indices, nodes, deep.materials))
            gpu.uploadSceneData(fine)
        cam = look_at((30.0, 0.4, 40.0), (30.0, 0.4, 0.0), fov=80.0)
        light = dict(pos=(30.0, 0.4, 10.0), N=(0.0, 0.0, -1.0), right=(1.0, 0.0, 0.0), up=(0.0, 1.0, 0.0), size=(20.0, 0.5), E=(5.0, 5.0, 5.0))
        params = make_params(64, 8, cam, fine.world_radius, len(fine.tris), light=light, max_bounces=2)
        run_lockstep(gpu, oracle_ctx(256), fine, params, iterations=4)


def test_pinned_host_memory_gives_the_same_bytes():
    """flx_host_alloc: uploading from and reading into page-locked arrays (the DMA path bench.py's e2e leg uses) gives exactly what
    the pageable path gives."""
    from fluctus_b200 import pinned_empty
    scene = make_room_scene(materials="mixed", textured=True)
    W, H, N = 64, 40, 2048
    params = room_params(scene, W, H, max_bounces=3)
    with CLContext(N) as a, CLContext(N) as b:
        ta, tb = setup_context(a, scene, params), setup_context(b, scene.pinned(), params)
        ta.start(); tb.start()
        for _ in range(5):
            ta.iterate(); tb.iterate()
        out = pinned_empty((W * H, 4), np.float32)
        got = b.readPixels(out)
        assert got is out
        compare_tasks(a.readTasks(), b.readTasks(), "pinned vs pageable upload")
        compare_pixels(a.readPixels(), out, "pinned vs pageable read-back", rtol=1e-5)


@pytest.mark.parametrize("direct", [1, 0])
def test_gather_on_its_own_stream_single_rank(tmp_path, direct):
    """flx_gather_pixels with a one-rank communicator (NCCL send/recv to self): the gather runs on the library's gather stream from
    a snapshot, so (1) the frame it delivers is the accumulator AS OF THE CALL even though rendering continues right behind it,
    (2) gathers every iteration do not disturb the render (same path state as without), (3) a resize between gathers gets fresh
    buffers of the right size (the full-image buffer kept its old size)."""
    scene = make_room_scene(materials="mixed")
    N = 4096
    with CLContext(N) as a, CLContext(N) as b:
        try:
            uid = a.commUniqueId()
            a.setTuning(gather_direct=direct)
            a.setTile(0, 1, 8)
            a.commInit(uid, 0, 1)
        except FluctusError as e:
            pytest.skip("NCCL not usable here: %s" % e)
        for (W, H) in ((64, 9), (64, 16), (48, 8)):
            params = room_params(scene, W, H, max_bounces=3)
            ta, tb = setup_context(a, scene, params), setup_context(b, scene, params)
            ta.start(); tb.start()
            frames = []
            for it in range(6):
                a.render(1); b.render(1)
                want = b.readPixels()
                full = np.zeros((W * H, 4), np.float32)
                if it % 2 == 0:
                    a.gatherPixels(0, full)          # blocking form: host image on return
                    frames.append((want, full))
                else:
                    a.gatherPixels(0)                # asynchronous form: rendering continues behind it
                    a.render(1); b.render(1)
            a.finishQueue()
            for want, full in frames:
                compare_pixels(full, want, "gathered frame %dx%d" % (W, H), rtol=1e-5)
            compare_tasks(a.readTasks(), b.readTasks(), "render with gathers vs without (%dx%d)" % (W, H))
            g_ms, g_n = a.checkTracingPerf()["gather"]
            assert g_n >= 6


def test_thin_lens_and_the_pinhole_shortcut():
    """Depth of field (wf_raygen.cl:59-63, mk_raygen.cl:49-53): with an open aperture the disk sample moves the ray origin; with aperture 0
    the offset is exactly +-0 and the library skips its cos / sin (lens_offset, flx_kernels.cuh) -- except when a component of the
    camera position is -0.0f, where origin + (+0) differs from origin in the sign bit and the full arithmetic must run.  All three
    cases in lockstep with the reference kernels, both integrators."""
    from parity_util import run_mk_lockstep
    scene = make_room_scene(materials="mixed", textured=True)
    W, H, N = 48, 32, 48 * 32
    for what, pos, aperture in (("open aperture", (0.0, 1.0, 0.95), 0.02), ("pinhole, -0.0 in the camera position", (-0.0, 1.0, 0.95), 0.0), ("pinhole", (0.1, 1.0, 0.95), 0.0)):
        cam = look_at(pos, (0.0, 0.9, -0.2), fov=70.0, aperture=aperture, focal_dist=1.2)
        light = dict(pos=(0.0, 1.98, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(0.3, 0.3), E=(60.0, 60.0, 60.0))
        params = make_params(W, H, cam, scene.world_radius, n_tris=len(scene.tris), light=light, max_bounces=3)
        with CLContext(N) as gpu:
            tg, _ = run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=8)
            orig = gpu.readTasks()[SLOT.ORIG:SLOT.ORIG + 3].view(np.float32)
        if aperture:
            assert len(np.unique(orig[0])) > 100, "an open aperture must spread the ray origins"
        with CLContext(N) as gpu:
            run_mk_lockstep(gpu, oracle_ctx(N), scene, params, spp=2)
