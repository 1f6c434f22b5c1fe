"""GPU parity: the CUDA path (through the C ABI) against the reference's own kernels compiled for the host
(oracle/_ref), on identical seeds -- complete path state bit-for-bit after every stage group, queue membership, raygen
queue order, per-pixel sample counts exactly, radiance bit-exact (or within 1e-4 relative where several paths may
splat one pixel in the same iteration and the float-atomic order is free)."""
import numpy as np
import pytest

from fluctus_b200 import CLContext, EnvMapData, SceneData
from fluctus_b200.scene import make_room_scene, room_params

from conftest import scene_blob
from parity_util import run_lockstep

pytestmark = pytest.mark.gpu


def oracle_ctx(n):
    from oracle.oracle_host import RefContext, PortContext, ref_available, port_available
    if ref_available():
        return RefContext(n)
    if port_available():
        return PortContext(n)
    pytest.skip("no oracle library built")


def synthetic_env(w=32, h=16, seed=3):
    rng = np.random.default_rng(seed)
    rgb = rng.uniform(0.0, 0.4, size=(h, w, 3)).astype(np.float32)
    rgb[3:5, 5:8] += 25.0  # a "sun"
    return EnvMapData.from_rgb(rgb)


@pytest.mark.parametrize("separate", [False, True])
def test_room_diffuse_area_light(separate):
    scene = make_room_scene(materials="diffuse")
    W, H, N = 96, 64, 4096
    params = room_params(scene, W, H, max_bounces=4, separate_queues=separate)
    with CLContext(N) as gpu:
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=16)


@pytest.mark.parametrize("separate", [False, True])
def test_room_all_bsdfs_textures_normal_map(separate):
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 96, 64, 6144
    params = room_params(scene, W, H, max_bounces=6, separate_queues=separate)
    with CLContext(N) as gpu:
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=20)


@pytest.mark.parametrize("area", [False, True])
def test_room_env_map_mis(area):
    scene = make_room_scene(materials="mixed", textured=True)
    # open the room: drop the ceiling and front wall so paths escape to the environment
    keep = np.ones(len(scene.tris), bool)
    keep[2:4] = False
    keep[6:8] = False
    from fluctus_b200.scene import build_bvh
    tris = scene.tris[keep]
    nodes, indices = build_bvh(tris)
    scene = SceneData(tris, indices, nodes, scene.materials, scene.tex_desc, scene.tex_data)
    W, H, N = 80, 48, 80 * 48
    params = room_params(scene, W, H, max_bounces=5, separate_queues=True, use_env_map=True, use_area_light=area, env_map_strength=2.0)
    with CLContext(N) as gpu:
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=16, env=synthetic_env())


@pytest.mark.parametrize("impl,expl", [(True, False), (False, True)])
def test_room_sampling_modes(impl, expl):
    scene = make_room_scene(materials="mixed")
    W, H, N = 64, 48, 64 * 48
    params = room_params(scene, W, H, max_bounces=4, sample_impl=impl, sample_expl=expl)
    with CLContext(N) as gpu:
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=10)


def test_room_russian_roulette_more_paths_than_pixels():
    scene = make_room_scene(materials="mixed")
    W, H, N = 48, 32, 4096  # N > W*H: several paths per pixel in flight, splat order free -> tolerance on RGB
    params = room_params(scene, W, H, max_bounces=3, use_roulette=True)
    with CLContext(N) as gpu:
        run_lockstep(gpu, oracle_ctx(N), scene, params, iterations=12)


def test_teapot_c1():
    """BASELINE config C1: teapot.ply, 2 bounces, Lambert, default camera and light (reduced to 128x128 for run time)."""
    scene = SceneData.load_blob(scene_blob("teapot"))
    from fluctus_b200 import make_params
    cam = dict(pos=(0, 1, 3.5), dir=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0), fov=60.0)
    W = H = 128
    params = make_params(W, H, cam, scene.world_radius, len(scene.tris), max_bounces=2)
    with CLContext(W * H) as gpu:
        run_lockstep(gpu, oracle_ctx(W * H), scene, params, iterations=16)


def test_conference_c2_small():
    """BASELINE config C2 (conference, 8 bounces, ceiling light) at 160x90 so the serial oracle finishes in seconds."""
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    W, H = 160, 90
    params = conference_params(scene, W, H)
    with CLContext(W * H) as gpu:
        run_lockstep(gpu, oracle_ctx(W * H), scene, params, iterations=12, check_every=3)


def test_stripe_tiling_with_one_part_is_the_identity():
    """flx_set_tile(0, 1, s) must not change anything: same path state and image as the untiled context."""
    scene = make_room_scene(materials="mixed")
    W, H, N = 64, 40, 2048
    params = room_params(scene, W, H, max_bounces=3)
    from parity_util import setup_context, compare_tasks, compare_pixels
    with CLContext(N) as a, CLContext(N) as b:
        b.setTile(0, 1, 8)
        ta, tb = setup_context(a, scene, params), setup_context(b, scene, params)
        ta.start(); tb.start()
        for _ in range(6):
            ta.iterate(); tb.iterate()
        compare_tasks(a.readTasks(), b.readTasks(), "tile(0,1,8) vs untiled")
        compare_pixels(a.readPixels(), b.readPixels(), "tile(0,1,8) vs untiled", rtol=1e-5)


def test_fused_render_equals_per_stage_loop():
    """flx_render (device-side counters, pixel index and stats) == the per-stage loop with host round trips."""
    scene = make_room_scene(materials="mixed", textured=True)
    W, H, N = 64, 40, 3000
    params = room_params(scene, W, H, max_bounces=4, separate_queues=True)
    from parity_util import setup_context, compare_tasks, compare_pixels
    with CLContext(N) as a, CLContext(N) as b:
        ta, tb = setup_context(a, scene, params), setup_context(b, scene, params)
        ta.start(); tb.start()
        for _ in range(9):
            ta.iterate()
        b.resetStats()
        tb.render(9)
        compare_tasks(a.readTasks(), b.readTasks(), "fused vs per-stage")
        compare_pixels(a.readPixels(), b.readPixels(), "fused vs per-stage", rtol=1e-5)
        st = b.getStats()
        assert (st.extensionRays, st.shadowRays, st.primaryRays, st.iterations) == (ta.stats["extensionRays"], ta.stats["shadowRays"], ta.stats["primaryRays"], 9)


@pytest.mark.parametrize("name", ["room_mixed_separate", "room_env_mis", "teapot_c1"])
def test_gpu_matches_golden(name):
    """The committed golden vectors (reference kernels, tests/golden/make_golden.py) reproduced bit-for-bit on the GPU."""
    import os
    from golden.make_golden import build_case
    from parity_util import setup_context, compare_tasks, compare_pixels
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    scene, params, env, n_tasks, iters = build_case(name)
    with CLContext(n_tasks) as gpu:
        tr = setup_context(gpu, scene, params, env)
        tr.start()
        compare_tasks(gpu.readTasks(), z["tasks_start"], "%s after prologue" % name)
        for _ in range(iters):
            tr.iterate()
        compare_tasks(gpu.readTasks(), z["tasks_end"], "%s after %d iterations" % (name, iters))
        compare_pixels(gpu.readPixels(), z["pixels"], name, rtol=1e-5)
        compare_pixels(gpu.readPreview(), z["preview"], name + " preview", rtol=1e-5)
        assert [tr.stats[k] for k in ("primaryRays", "extensionRays", "shadowRays")] == list(z["stats"])


def test_traversal_work_counters_match_the_oracle():
    """flx_set_counting: V/B/T/U per ray counted on the GPU == counted by the instrumented C restatement."""
    import ctypes as C
    from oracle.oracle_host import PortContext, port_available
    if not port_available():
        pytest.skip("oracle/liboracle.so not built")
    from parity_util import setup_context
    scene = make_room_scene(materials="diffuse")
    W, H, N = 64, 40, 2560
    params = room_params(scene, W, H, max_bounces=4)
    cpu = PortContext(N)
    with CLContext(N) as gpu:
        tg, tc = setup_context(gpu, scene, params), setup_context(cpu, scene, params)
        tg.start(); tc.start()
        gpu.setCounting(True)
        cpu.lib.port_count_work(1)
        for _ in range(5):
            tg.iterate(); tc.iterate()
        e, s = (C.c_ulonglong * 5)(), (C.c_ulonglong * 5)()
        cpu.lib.port_work_counts(e, s)
        cpu.lib.port_count_work(0)
        g = gpu.getTraceCounts()
        assert [g["ext"][k] for k in ("nodes", "boxes", "tris", "updates", "rays")] == list(e)
        assert [g["shadow"][k] for k in ("nodes", "boxes", "tris")] == list(s)[:3]


@pytest.mark.parametrize("direct", [1, 0])
def test_multi_gpu_tiles_and_nccl_gather(direct):
    import subprocess, sys, os, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29731 + direct), os.path.join(root, "tests", "multi_gpu_check.py")], capture_output=True, text=True, timeout=600,
                       env=dict(os.environ, FLX_GATHER_DIRECT=str(direct)))
    assert r.returncode == 0 and "MULTI_GPU_OK" in r.stdout and "GATHERED_IMAGE_OK" in r.stdout and "RESIZE_GATHER_OK 64x16" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


def test_cpp_wrapper_headless_driver_matches_golden(tmp_path):
    """The C++ host side (include/fluctus_b200/clcontext.hpp, examples/flx_headless.cpp: the reference's benchmark loop with
    its call sites unchanged) reproduces the teapot golden vector produced by the reference kernels."""
    import json, os, subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "flx_headless")
    if not os.path.exists(exe):
        pytest.skip("examples/flx_headless not built (python __graft_entry__.py)")
    z = np.load(os.path.join(root, "tests", "golden", "teapot_c1.npz"))
    out = tmp_path / "pix.rgba"
    r = subprocess.run([exe, scene_blob("teapot"), "64", "64", "4096", "2", "16", str(out)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert [res["primary"], res["extension"], res["shadow"]] == list(z["stats"])
    pix = np.fromfile(out, np.float32).reshape(-1, 4)
    from parity_util import compare_pixels
    compare_pixels(pix, z["pixels"], "C++ driver vs golden", rtol=1e-5)


def test_luxball_c4_small():
    """BASELINE config C4 (luxball, ideal dielectric, 16 bounces, separate material queues) at 96x54."""
    scene = SceneData.load_blob(scene_blob("luxball"))
    from bench_configs import luxball_params
    W, H = 96, 54
    params = luxball_params(scene, W, H)
    with CLContext(W * H) as gpu:
        run_lockstep(gpu, oracle_ctx(W * H), scene, params, iterations=20, check_every=4)


def test_country_kitchen_c3_small():
    """BASELINE config C3 (country kitchen: glossy / GGX / mirror / dielectric materials, 17 textures + a bump map,
    night.hdr environment map with alias-method IBL, MIS, separate queues) at 96x54."""
    scene = SceneData.load_blob(scene_blob("country_kitchen"))
    import os
    from conftest import SCENES_DIR
    envp = os.path.join(SCENES_DIR, "night.env.bin")
    if not os.path.exists(envp):
        pytest.skip("env map blob missing")
    env = EnvMapData.load_blob(envp)
    from bench_configs import kitchen_params
    W, H = 96, 54
    params = kitchen_params(scene, W, H)
    with CLContext(W * H) as gpu:
        run_lockstep(gpu, oracle_ctx(W * H), scene, params, iterations=12, env=env, check_every=4)


def test_country_kitchen_c3_from_files_only():
    """C3 with no blob, no reference code and no Pillow on the way in: Country-Kitchen.obj + .mtl through flx_scene_load, its JPEG
    and PNG textures through flx_image_load + flx_pack_textures, night.hdr through flx_envmap_load, the hierarchy from the GPU
    builder -- rendered in lockstep with the reference's kernels on the same arrays."""
    import os
    from fluctus_b200.scene_io import load_envmap, load_model_with_textures
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "assets")
    obj, hdr = os.path.join(root, "country_kitchen", "Country-Kitchen.obj"), os.path.join(root, "env_maps", "night.hdr")
    if not (os.path.exists(obj) and os.path.exists(hdr)):
        pytest.skip("asset files not under oracle/_ref/assets (oracle/make_scenes.py copies them where /root/reference exists)")
    model, desc, data = load_model_with_textures(obj)
    env = load_envmap(hdr)
    from bench_configs import kitchen_params
    W, H = 96, 54
    with CLContext(W * H) as gpu:
        nodes, idx, ms = gpu.buildBVH(model.tris, 8, "ploc")
        scene = SceneData(model.tris, idx, nodes, model.materials, desc, data)
        params = kitchen_params(scene, W, H)
        tg, tc = run_lockstep(gpu, oracle_ctx(W * H), scene, params, iterations=12, env=env, check_every=4)
        assert tg.stats["extensionRays"] > 0 and len(model.texture_names) == 11
