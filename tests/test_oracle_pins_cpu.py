"""Pins of the ORACLE itself (the reference has no tests or golden vectors for this path, so the oracle -- the
reference's kernels compiled for the host -- is checked for self-consistency instead):

  * the stackless bit-stack traversal (-DUSE_BITSTACK, src/bvh.cl:10-230) finds the same hits as the explicit-stack one;
  * the array-of-structures build (no -DUSE_SOA, src/geom.h:26-37) computes the same path state as the SoA build;
  * one material queue vs per-type queues (wfSeparateQueues) give the same path state;
  * a white furnace with a closed form, under BSDF sampling, light sampling and MIS;
  * the C library's sinf / cosf / acosf / atan2f / powf instead of include/flx_math.h give the same picture within Monte-Carlo
    error -- the one check that can see a bias in the math header the GPU, the restatement and the base oracle all share.
Host code only."""
import numpy as np
import pytest

from fluctus_b200 import SceneData
from fluctus_b200.scene import make_room_scene, room_params

from conftest import scene_blob
from parity_util import compare_tasks, run_lockstep, setup_context

from oracle.oracle_host import RefContext, ref_available

pytestmark = pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("variant", ["bs_", "aos_"])
def test_reference_build_variants_agree_on_a_room_with_every_bsdf(variant):
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 48, 32, 2048
    for separate in (False, True):
        params = room_params(scene, W, H, max_bounces=5, separate_queues=separate)
        run_lockstep(RefContext(N, variant=variant), RefContext(N), scene, params, iterations=10, exact_rgb=True)


@pytest.mark.parametrize("variant", ["bs_", "aos_"])
def test_reference_build_variants_agree_on_conference(variant):
    """the reference's own SBVH (duplicated references, leaves of up to 8) through both traversals / both layouts"""
    scene = SceneData.load_blob(scene_blob("conference"))
    from bench_configs import conference_params
    W, H = 96, 54
    params = conference_params(scene, W, H)
    run_lockstep(RefContext(W * H, variant=variant), RefContext(W * H), scene, params, iterations=10, check_every=3, exact_rgb=True)


def test_single_and_separate_material_queues_compute_the_same_paths():
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 48, 32, 48 * 32
    a, b = RefContext(N), RefContext(N)
    ta = setup_context(a, scene, room_params(scene, W, H, max_bounces=5, separate_queues=False))
    tb = setup_context(b, scene, room_params(scene, W, H, max_bounces=5, separate_queues=True))
    ta.start()
    tb.start()
    for it in range(10):
        ca, cb = ta.iterate(), tb.iterate()
        assert (ca.raygenQueue, ca.extensionQueue, ca.shadowQueue) == (cb.raygenQueue, cb.extensionQueue, cb.shadowQueue)
        assert ca.diffuseQueue == cb.diffuseQueue + cb.glossyQueue + cb.ggxReflQueue + cb.ggxRefrQueue + cb.deltaQueue
        compare_tasks(a.readTasks(), b.readTasks(), "single vs separate queues, iteration %d" % it)
    assert np.array_equal(a.readPixels().view(np.uint32), b.readPixels().view(np.uint32))


def test_white_furnace_closed_form_and_estimators_agree():
    from furnace_util import check_closed_form
    check_closed_form(lambda n: RefContext(n, parallel_trace=True), width=40, height=30, iterations=200)


def test_libm_math_oracle_agrees_within_monte_carlo_error():
    """Same seeds, same kernels; only sin / cos / tan / acos / atan2 / pow come from glibc instead of include/flx_math.h.  A last-bit
    difference in a sampled direction is harmless until it flips a comparison (a different triangle, a different alias-table
    column) -- from there that path is an independent sample.  So: most pixels agree to many digits, the rest within noise, and
    the image means agree far inside the noise of either."""
    from furnace_util import render
    from fluctus_b200 import EnvMapData
    from test_oracle_cpu import open_room
    scene = open_room()  # no ceiling, no front wall: paths escape to the environment, so acos / atan2 (direction -> lat-long) take part
    W, H, N = 40, 30, 1200
    rng = np.random.default_rng(5)
    env = EnvMapData.from_rgb(rng.uniform(0.1, 0.5, size=(16, 32, 3)).astype(np.float32))
    imgs = []
    for variant in ("", "lm_"):
        ctx = RefContext(N, variant=variant, parallel_trace=True)
        params = room_params(scene, W, H, max_bounces=5, separate_queues=True, use_env_map=True, use_area_light=True, env_map_strength=1.5)
        img, cnt = render(ctx, scene, params, env, 1600)
        imgs.append((img.astype(np.float64), cnt))
    (a, ca), (b, cb) = imgs
    assert abs(ca.sum() - cb.sum()) < 0.01 * ca.sum(), "the two builds trace a different number of paths"
    mean_rel = np.abs(a.mean(axis=0) - b.mean(axis=0)) / a.mean(axis=0)
    assert (mean_rel < 0.02).all(), "image means differ by %r: a bias in one of the two math libraries" % (mean_rel,)
    # block means (5x5 pixels): a bias confined to one material or one light would show up here
    def blocks(x):
        return x.reshape(H // 5, 5, W // 5, 5, 3).mean(axis=(1, 3))
    blk = np.abs(blocks(a) - blocks(b)) / np.maximum(blocks(a), 1e-3)
    assert np.median(blk) < 0.03 and blk.max() < 0.5, (float(np.median(blk)), float(blk.max()))
