"""CPU tests of the oracle itself (no GPU): the C restatement (oracle/wf_oracle.c) against the reference's own kernels
compiled for the host (oracle/_ref, when built) in lockstep, and against the committed golden vectors that those
kernels produced (tests/golden/, always)."""
import os

import numpy as np
import pytest

from fluctus_b200 import EnvMapData, SceneData, Tracer, make_params
from fluctus_b200.scene import build_bvh, make_room_scene, room_params

from conftest import scene_blob
from parity_util import run_lockstep, run_mk_lockstep, setup_context, compare_tasks, compare_mk_tasks, compare_pixels

from oracle.oracle_host import PortContext, RefContext, port_available, ref_available

needs_ref = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
needs_port = pytest.mark.skipif(not port_available(), reason="oracle/liboracle.so not built (python oracle/build_oracle.py)")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def synthetic_env(w=32, h=16, seed=3):
    rng = np.random.default_rng(seed)
    rgb = rng.uniform(0.0, 0.4, size=(h, w, 3)).astype(np.float32)
    rgb[3:5, 5:8] += 25.0
    return EnvMapData.from_rgb(rgb)


def open_room(textured=True):
    scene = make_room_scene(materials="mixed", textured=textured)
    keep = np.ones(len(scene.tris), bool)
    keep[2:4] = False
    keep[6:8] = False
    tris = scene.tris[keep]
    nodes, indices = build_bvh(tris)
    return SceneData(tris, indices, nodes, scene.materials, scene.tex_desc, scene.tex_data)


@needs_ref
@needs_port
@pytest.mark.parametrize("separate", [False, True])
def test_port_matches_reference_kernels_all_bsdfs(separate):
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 64, 48, 4096
    params = room_params(scene, W, H, max_bounces=6, separate_queues=separate)
    run_lockstep(PortContext(N), RefContext(N), scene, params, iterations=14, exact_rgb=True)


@needs_ref
@needs_port
@pytest.mark.parametrize("area", [False, True])
def test_port_matches_reference_kernels_env_map(area):
    scene = open_room()
    W, H = 48, 32
    params = room_params(scene, W, H, max_bounces=5, separate_queues=True, use_env_map=True, use_area_light=area, env_map_strength=2.0)
    run_lockstep(PortContext(W * H), RefContext(W * H), scene, params, iterations=12, env=synthetic_env(), exact_rgb=True)


@needs_ref
@needs_port
def test_port_matches_reference_kernels_roulette_and_sampling_modes():
    scene = make_room_scene(materials="mixed")
    for impl, expl, rr in ((True, False, False), (False, True, False), (True, True, True)):
        params = room_params(scene, 40, 30, max_bounces=3, sample_impl=impl, sample_expl=expl, use_roulette=rr)
        run_lockstep(PortContext(2048), RefContext(2048), scene, params, iterations=8, exact_rgb=True)


@needs_ref
@needs_port
def test_port_matches_reference_kernels_conference():
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    params = conference_params(scene, 96, 54)
    run_lockstep(PortContext(96 * 54), RefContext(96 * 54), scene, params, iterations=10, check_every=3, exact_rgb=True)


@needs_ref
@needs_port
def test_work_counters_match_survey_figures():
    """The instrumented restatement reproduces the per-ray traversal work the survey measured on the reference
    (~24 node pops, ~44 box tests, ~7 triangle tests per extension ray on conference)."""
    import ctypes as C
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    params = conference_params(scene, 96, 54)
    ctx = PortContext(96 * 54)
    tr = setup_context(ctx, scene, params)
    tr.start()
    ctx.lib.port_count_work(1)
    for _ in range(10):
        tr.iterate()
    e, s = (C.c_ulonglong * 5)(), (C.c_ulonglong * 5)()
    ctx.lib.port_work_counts(e, s)
    ctx.lib.port_count_work(0)
    V, B, T, U = (e[k] / e[4] for k in range(4))
    assert 15 < V < 35 and 30 < B < 60 and 4 < T < 12 and 0.8 < U < 1.6, (V, B, T, U)
    assert e[1] % 2 == 0  # two child boxes per inner-node pop (bvh.cl:283-284)
    assert 10 < s[0] / s[4] < 30


@needs_port
@pytest.mark.parametrize("name", ["room_mixed_separate", "room_env_mis", "teapot_c1"])
def test_port_matches_golden(name):
    """Golden vectors = path state and radiance written by the reference's kernels (tests/golden/make_golden.py)."""
    from golden.make_golden import CASES, build_case
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.fail("golden fixture %s is missing" % path)
    z = np.load(path)
    scene, params, env, n_tasks, iters = build_case(name, blob_loader=scene_blob)
    ctx = PortContext(n_tasks)
    tr = setup_context(ctx, scene, params, env)
    tr.start()
    compare_tasks(ctx.readTasks(), z["tasks_start"], "%s after prologue" % name)
    for _ in range(iters):
        tr.iterate()
    compare_tasks(ctx.readTasks(), z["tasks_end"], "%s after %d iterations" % (name, iters))
    compare_pixels(ctx.readPixels(), z["pixels"], name, exact_rgb=True)
    compare_pixels(ctx.readPreview(), z["preview"], name + " preview", exact_rgb=True)
    assert [tr.stats[k] for k in ("primaryRays", "extensionRays", "shadowRays")] == list(z["stats"])


# ---------------------------------------------------------------------------------------------- microkernel integrator (src/mk_*.cl)
@needs_ref
@needs_port
def test_port_mk_matches_reference_kernels_all_bsdfs():
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H = 48, 32
    params = room_params(scene, W, H, max_bounces=5)
    run_mk_lockstep(PortContext(W * H), RefContext(W * H), scene, params, spp=4)


@needs_ref
@needs_port
@pytest.mark.parametrize("area", [False, True])
def test_port_mk_matches_reference_kernels_env_map(area):
    scene = open_room()
    W, H = 40, 24
    params = room_params(scene, W, H, max_bounces=4, use_env_map=True, use_area_light=area, env_map_strength=2.0)
    run_mk_lockstep(PortContext(W * H), RefContext(W * H), scene, params, spp=3, env=synthetic_env())


@needs_ref
@needs_port
def test_port_mk_matches_reference_kernels_interactive_roulette_sampling_modes_and_short_task_buffer():
    """Tracer::update's preview + progressive calls (src/tracer.cpp:267-299); NUM_TASKS < width*height renders only the first
    NUM_TASKS pixels (limit = min(width*height, numTasks), e.g. src/mk_raygen.cl:9)."""
    scene = make_room_scene(materials="mixed")
    for impl, expl, rr, n in ((True, False, False, 1200), (False, True, False, 700), (True, True, True, 1000)):
        params = room_params(scene, 40, 30, max_bounces=3, sample_impl=impl, sample_expl=expl, use_roulette=rr)
        run_mk_lockstep(PortContext(n), RefContext(n), scene, params, spp=6, interactive=True)


@needs_ref
@needs_port
def test_port_mk_matches_reference_kernels_conference():
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    params = conference_params(scene, 64, 36)
    params.maxBounces = 4
    run_mk_lockstep(PortContext(64 * 36), RefContext(64 * 36), scene, params, spp=2, check_every=2)


@needs_port
@pytest.mark.parametrize("name", ["mk_room_env_mis", "mk_room_mixed"])
def test_port_mk_matches_golden(name):
    """Golden vectors written by the reference's own mk_*.cl kernels compiled for the host (tests/golden/make_golden.py)."""
    from golden.make_golden import MK_CASES, build_case, run_mk_case
    path = os.path.join(GOLDEN, name + ".npz")
    if not os.path.exists(path):
        pytest.fail("golden fixture %s is missing" % path)
    z = np.load(path)
    out = run_mk_case(PortContext(build_case(MK_CASES[name][0], scene_blob)[3]), name, scene_blob)
    n_live = int(z["n_live"][0])
    compare_mk_tasks(out["tasks_first_bounce"], z["tasks_first_bounce"], name + " after the first bounce", n_live)
    compare_mk_tasks(out["tasks_end"], z["tasks_end"], name + " at the end", n_live)
    compare_pixels(out["pixels"], z["pixels"], name, exact_rgb=True)
    compare_pixels(out["preview"], z["preview"], name + " preview", exact_rgb=True)
    assert list(out["stats"]) == list(z["stats"])


# ---------------------------------------------------------------------------------------------- the other BASELINE scenes, small
def _kitchen():
    scene = SceneData.load_blob(scene_blob("country_kitchen"))
    envp = os.path.join(os.path.dirname(scene_blob("country_kitchen")), "night.env.bin")
    if not os.path.exists(envp):
        pytest.skip("env map blob missing")
    from bench_configs import kitchen_params
    return scene, EnvMapData.load_blob(envp), kitchen_params


@needs_ref
@needs_port
def test_port_matches_reference_kernels_country_kitchen_and_luxball():
    """C3 (GGX / glossy / mirror / dielectric materials, 11 textures, a bump map, night.hdr alias-method IBL, separate queues) and
    C4 (ideal dielectric, 16 bounces) at thumbnail size: restatement == the reference's kernels, wavefront integrator."""
    scene, env, kitchen_params = _kitchen()
    W, H = 48, 27
    run_lockstep(PortContext(W * H), RefContext(W * H), scene, kitchen_params(scene, W, H), iterations=10, env=env, check_every=3, exact_rgb=True)
    from bench_configs import luxball_params
    lux = SceneData.load_blob(scene_blob("luxball"))
    run_lockstep(PortContext(W * H), RefContext(W * H), lux, luxball_params(lux, W, H), iterations=18, check_every=6, exact_rgb=True)


@needs_ref
@needs_port
def test_port_mk_matches_reference_kernels_country_kitchen():
    """the microkernel integrator on C3: every BSDF type, textures, env-map + MIS in one sampleBsdf kernel"""
    scene, env, kitchen_params = _kitchen()
    W, H = 40, 24
    params = kitchen_params(scene, W, H, max_bounces=4)
    run_mk_lockstep(PortContext(W * H), RefContext(W * H), scene, params, spp=2, env=env, check_every=2)


@needs_ref
def test_parallel_trace_oracle_is_the_serial_oracle():
    """The oracle mode the full-size GPU parity tests use (tests/test_gpu_parity_large.py): traversal kernels through the
    OpenMP build, everything that decides queue order serial.  Must be bit-identical to the all-serial oracle."""
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    W, H = 128, 72
    params = conference_params(scene, W, H)
    run_lockstep(RefContext(W * H, parallel_trace=True), RefContext(W * H), scene, params, iterations=8, check_every=4, exact_rgb=True)
    if port_available():
        run_lockstep(PortContext(W * H, parallel_trace=True), RefContext(W * H), scene, params, iterations=4, check_every=4, exact_rgb=True)
