"""GPU parity of the microkernel integrator (reference src/mk_*.cl, driven like Tracer::renderSingle,
src/tracer.cpp:95-169, and the non-wavefront branch of Tracer::update, src/tracer.cpp:267-299): the CUDA path through the
C ABI against the reference's own mk kernels compiled for the host, kernel by kernel -- complete path state including the
phase word bit-for-bit, accumulator and preview bit-for-bit (path g owns pixel g, so there are no float atomics), ray and
sample statistics exactly; then against the committed golden vectors, and at full size through size-independent properties."""
import os

import numpy as np
import pytest

from fluctus_b200 import CLContext, SceneData, Tracer
from fluctus_b200.scene import build_bvh, make_room_scene, room_params

from conftest import scene_blob
from parity_util import compare_mk_tasks, compare_pixels, mk_stats, run_mk_lockstep, setup_context
from test_gpu_parity import oracle_ctx, synthetic_env

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def open_room():
    scene = make_room_scene(materials="mixed", textured=True)
    keep = np.ones(len(scene.tris), bool)
    keep[2:4] = False
    keep[6:8] = False
    tris = scene.tris[keep]
    nodes, indices = build_bvh(tris)
    return SceneData(tris, indices, nodes, scene.materials, scene.tex_desc, scene.tex_data)


def test_mk_room_all_bsdfs_textures_normal_map():
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H = 96, 64
    params = room_params(scene, W, H, max_bounces=6)
    with CLContext(W * H) as gpu:
        run_mk_lockstep(gpu, oracle_ctx(W * H), scene, params, spp=3)


@pytest.mark.parametrize("area", [False, True])
def test_mk_room_env_map_mis(area):
    """both light samples of one vertex in flight at once: env-map ray (the light quad blocks it) + area-light ray"""
    scene = open_room()
    W, H = 80, 48
    params = room_params(scene, W, H, max_bounces=5, use_env_map=True, use_area_light=area, env_map_strength=2.0)
    with CLContext(W * H) as gpu:
        run_mk_lockstep(gpu, oracle_ctx(W * H), scene, params, spp=3, env=synthetic_env())


@pytest.mark.parametrize("impl,expl,rr,n", [(True, False, False, 64 * 48 + 500), (False, True, False, 1777), (True, True, True, 64 * 48)])
def test_mk_interactive_loop_sampling_modes_roulette_ragged_sizes(impl, expl, rr, n):
    """Tracer::update's preview (two segments + splatPreview) then progressive calls; more tasks than pixels, fewer tasks
    than pixels (only the first NUM_TASKS pixels render, src/mk_raygen.cl:9) and sizes that are no multiple of the CTA."""
    scene = make_room_scene(materials="mixed")
    params = room_params(scene, 64, 48, max_bounces=3, sample_impl=impl, sample_expl=expl, use_roulette=rr)
    with CLContext(n) as gpu:
        run_mk_lockstep(gpu, oracle_ctx(n), scene, params, spp=8, interactive=True)


def test_mk_conference():
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    W, H = 128, 72
    params = conference_params(scene, W, H, max_bounces=4)
    with CLContext(W * H) as gpu:
        run_mk_lockstep(gpu, oracle_ctx(W * H), scene, params, spp=2, check_every=2)


@pytest.mark.parametrize("name", ["mk_room_env_mis", "mk_room_mixed"])
def test_mk_matches_golden(name):
    from golden.make_golden import MK_CASES, build_case, run_mk_case
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    n = build_case(MK_CASES[name][0], scene_blob)[3]
    with CLContext(n) as gpu:
        out = run_mk_case(gpu, name, scene_blob)
    n_live = int(z["n_live"][0])
    compare_mk_tasks(out["tasks_first_bounce"], z["tasks_first_bounce"], name + " after the first bounce", n_live)
    compare_mk_tasks(out["tasks_end"], z["tasks_end"], name + " at the end", n_live)
    compare_pixels(out["pixels"], z["pixels"], name, exact_rgb=True)
    compare_pixels(out["preview"], z["preview"], name + " preview", exact_rgb=True)
    assert list(out["stats"]) == list(z["stats"])


def test_mk_full_size_invariants_and_fused_loop():
    """Conference 1920x1080 (too large for the CPU oracle): (1) every pixel holds exactly spp samples -- the guarantee
    renderSingle switches integrators for (src/tracer.cpp:99); (2) the fused loop (flx_render_single) and the call-by-call
    loop give bit-identical accumulators and statistics; (3) the image agrees with the wavefront integrator's in the mean
    (same estimator, different sample sets)."""
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    W, H, spp = 1920, 1080, 4
    params = conference_params(scene, W, H, max_bounces=4)
    images = []
    for fused in (False, True):
        with CLContext(W * H) as gpu:
            tr = setup_context(gpu, scene, params)
            gpu.resetStats()
            tr.renderSingle(spp, fused=fused)
            pix = gpu.readPixels()
            images.append((pix, mk_stats(gpu)))
            assert np.array_equal(pix[:, 3], np.full(W * H, float(spp), np.float32))
            assert np.isfinite(pix).all()
    assert np.array_equal(images[0][0].view(np.uint32), images[1][0].view(np.uint32))
    assert images[0][1] == images[1][1]
    prim, ext, shadow, samples = images[0][1]
    assert prim == samples == W * H * spp and ext > 0 and shadow > 0
    with CLContext(1 << 21) as gpu:
        tr = setup_context(gpu, scene, params)
        tr.start()
        tr.render(100)
        wf = gpu.readPixels()
    mk_mean = (images[0][0][:, :3].sum(axis=0) / images[0][0][:, 3].sum()).astype(np.float64)
    wf_mean = (wf[:, :3].sum(axis=0) / wf[:, 3].sum()).astype(np.float64)
    assert np.allclose(mk_mean, wf_mean, rtol=0.05), (mk_mean, wf_mean)


def test_file_to_picture_pipeline_without_reference_code(tmp_path):
    """examples/flx_render_file: OBJ -> flx_scene_load -> flx_build_bvh (GPU) -> upload -> Tracer::renderSingle's loop ->
    display pass -> CLContext::saveImage, all through the C ABI / C++ wrapper.  The model is the procedural test room
    written out as an OBJ; the picture must decode, have the requested size and show the lit room."""
    import json
    import subprocess
    from PIL import Image
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "flx_render_file")
    if not os.path.exists(exe):
        pytest.skip("examples/flx_render_file not built (python __graft_entry__.py)")
    room = make_room_scene(materials="diffuse", n_blobs=8)
    obj = tmp_path / "room.obj"
    with open(obj, "w") as f:
        for t in room.tris:
            for v in ("v0", "v1", "v2"):
                f.write("v %.9g %.9g %.9g\n" % tuple(t[v]["p"][:3]))
        for i in range(len(room.tris)):
            f.write("f %d %d %d\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    png = tmp_path / "room.png"
    r = subprocess.run([exe, str(obj), str(png), "160", "96", "8", "3"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert res["triangles"] == len(room.tris) and res["spp"] == 8 and res["bvh_build_ms"] < 50
    img = np.asarray(Image.open(png).convert("RGB"))
    assert img.shape == (96, 160, 3) and img.mean() > 5 and img.std() > 2
    hdr = tmp_path / "room.hdr"
    r = subprocess.run([exe, str(obj), str(hdr), "160", "96", "4", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    from fluctus_b200.scene_io import load_envmap
    lin = load_envmap(hdr).rgb
    assert lin.shape == (96, 160, 3) and np.isfinite(lin).all() and lin.max() > 0
    r = subprocess.run([exe, str(tmp_path / "missing.obj"), str(png)], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "cannot open" in r.stderr
    # a textured model: the PNG texture is decoded and packed by the library, the (missing) JPEG falls back to the material constant
    tex = np.zeros((8, 8, 3), np.uint8)
    tex[..., 1] = 255  # pure green albedo
    Image.fromarray(tex, "RGB").save(tmp_path / "green.png")
    with open(tmp_path / "tex.mtl", "w") as f:
        f.write("newmtl painted\nKd 1 0 0\nmap_Kd green.png\nnewmtl photo\nKd 0.5 0.5 0.5\nmap_Kd missing.jpg\n")
    with open(tmp_path / "tex.obj", "w") as f:
        f.write("mtllib tex.mtl\n")
        for t in room.tris:
            for v in ("v0", "v1", "v2"):
                f.write("v %.9g %.9g %.9g\n" % tuple(t[v]["p"][:3]))
        f.write("vt 0.5 0.5\nusemtl painted\n")
        for i in range(len(room.tris)):
            if i == len(room.tris) // 2:
                f.write("usemtl photo\n")
            f.write("f %d/1 %d/1 %d/1\n" % (3 * i + 1, 3 * i + 2, 3 * i + 3))
    out = tmp_path / "tex.png"
    r = subprocess.run([exe, str(tmp_path / "tex.obj"), str(out), "96", "64", "8", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert "1 of 2 decoded" in r.stderr
    img = np.asarray(Image.open(out).convert("RGB")).astype(np.float64)
    assert img[..., 1].mean() > 1.5 * img[..., 0].mean()  # the green texture shows (the red Kd it overrides would not)


def test_mk_tiled_contexts_cover_the_image():
    """The microkernel integrator under flx_set_tile (image rows dealt to the parts in interleaved stripes): two
    contexts on one GPU render the two halves of the stripes; de-interleaved, every pixel holds exactly spp samples and the
    picture agrees row by row with the untiled render (different seed -> pixel map, same estimator; a wrong stripe map would
    permute rows)."""
    from fluctus_b200 import dist as fd
    scene = make_room_scene(materials="diffuse", n_blobs=8)
    W, H, S, spp = 64, 44, 4, 48  # 11 stripes of 4 rows: part 0 gets 6, part 1 gets 5
    params = room_params(scene, W, H, max_bounces=2)
    full = np.zeros((H, W, 4), np.float32)
    for part in range(2):
        with CLContext(W * H) as gpu:
            gpu.setTile(part, 2, S)
            tr = setup_context(gpu, scene, params)
            tr.renderSingle(spp, fused=True)
            tile = gpu.readPixels().reshape(-1, W, 4)
            rows = [y for y in range(H) if (y // S) % 2 == part]
            assert len(rows) == tile.shape[0] == fd.tile_pixels(W, H, part, 2, S) // W
            full[rows] = tile
    assert np.array_equal(full[..., 3], np.full((H, W), float(spp), np.float32))
    with CLContext(W * H) as gpu:
        tr = setup_context(gpu, scene, params)
        tr.renderSingle(spp, fused=True)
        solo = gpu.readPixels().reshape(H, W, 4)
    row_a, row_b = full[..., :3].mean(axis=(1, 2)), solo[..., :3].mean(axis=(1, 2))
    assert np.allclose(row_a, row_b, rtol=0.08, atol=1e-3), np.abs(row_a - row_b).max()
    assert abs(full[..., :3].mean() - solo[..., :3].mean()) < 0.02 * solo[..., :3].mean()


def test_mk_country_kitchen_c3_small():
    """BASELINE config C3 through the microkernel integrator at thumbnail size: every BSDF type on its own shading list, 11 textures
    and a bump map, night.hdr alias-method IBL with MIS -- kernel by kernel against the reference's mk kernels."""
    from fluctus_b200 import EnvMapData
    from conftest import SCENES_DIR
    scene = SceneData.load_blob(scene_blob("country_kitchen"))
    envp = os.path.join(SCENES_DIR, "night.env.bin")
    if not os.path.exists(envp):
        pytest.skip("env map blob missing")
    from bench_configs import kitchen_params
    W, H = 64, 36
    params = kitchen_params(scene, W, H, max_bounces=5)
    with CLContext(W * H) as gpu:
        run_mk_lockstep(gpu, oracle_ctx(W * H), scene, params, spp=2, env=EnvMapData.load_blob(envp), check_every=2)
