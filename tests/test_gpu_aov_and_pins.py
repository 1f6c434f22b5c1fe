"""GPU: the denoiser feature buffers (the reference's USE_OPTIX_DENOISER build, src/wf_logic.cl:186-209, src/mk_next_vertex.cl:60-70,
src/mk_sample_bsdf.cl:56-66, src/mk_postprocess.cl:49-54) against the reference kernels built that way (oracle variant "dn_"),
and the oracle's self-consistency pins (tests/test_oracle_pins_cpu.py) repeated on the CUDA path: the white furnace with its
closed form, and the C-library-math oracle."""
import numpy as np
import pytest

from fluctus_b200 import CLContext, EnvMapData
from fluctus_b200.scene import make_room_scene, room_params

from parity_util import compare_pixels, run_lockstep, run_mk_lockstep

pytestmark = pytest.mark.gpu


def dn_oracle(n):
    from oracle.oracle_host import RefContext, ref_available
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    return RefContext(n, variant="dn_")


def compare_aovs(gpu, cpu, what, exact):
    for which in ("normal", "albedo"):
        for processed in (False, True):
            a, b = gpu.readDenoiserAOV(which, processed), cpu.readDenoiserAOV(which, processed)
            compare_pixels(a, b, "%s: %s buffer%s" % (what, which, " after the display pass" if processed else ""), rtol=1e-5, exact_rgb=exact)
    n = gpu.readDenoiserAOV("normal")
    assert n[:, 3].sum() > 0 and np.abs(n[:, :3]).sum() > 0, "no normals were accumulated"
    assert gpu.readDenoiserAOV("albedo")[:, 3].sum() > 0, "no albedo was accumulated"


@pytest.mark.parametrize("separate", [False, True])
def test_wavefront_denoiser_buffers_match_the_reference_build(separate):
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H, N = 64, 48, 4096
    params = room_params(scene, W, H, max_bounces=5, separate_queues=separate)
    cpu = dn_oracle(N)
    with CLContext(N) as gpu:
        gpu.setDenoiser(True)
        run_lockstep(gpu, cpu, scene, params, iterations=14)  # path state incl. firstDiffuseHit, accumulator, preview
        compare_aovs(gpu, cpu, "wavefront", exact=False)      # several paths may add to one pixel: float-atomic order is free
        # and through the fused loop
        from fluctus_b200 import Tracer
        gpu.render(5)
        tr = Tracer(cpu, params)
        for _ in range(5):
            tr.iterate()
        gpu.enqueuePostprocessKernel(params)
        gpu.finishQueue()
        compare_aovs(gpu, cpu, "wavefront, fused loop", exact=False)


def test_denoiser_buffers_are_off_by_default_and_cost_nothing_in_the_state():
    scene = make_room_scene(materials="mixed")
    W, H, N = 48, 32, 2048
    params = room_params(scene, W, H, max_bounces=4)
    from oracle.oracle_host import RefContext
    with CLContext(N) as gpu:
        run_lockstep(gpu, RefContext(N), scene, params, iterations=8)
        a = gpu.readDenoiserAOV("albedo")
        assert np.array_equal(a, np.tile(np.float32([0.1, 0.1, 0.1, 0.0]), (W * H, 1))), "albedo buffer must stay at its reset value (wf_reset.cl:22)"
        assert not gpu.readDenoiserAOV("normal").any()


def test_microkernel_denoiser_buffers_match_the_reference_build():
    scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
    W, H = 48, 36
    N = W * H
    params = room_params(scene, W, H, max_bounces=4)
    cpu = dn_oracle(N)
    with CLContext(N) as gpu:
        gpu.setDenoiser(True)
        run_mk_lockstep(gpu, cpu, scene, params, spp=3)
        compare_aovs(gpu, cpu, "microkernel", exact=True)  # path g owns pixel g: no atomics


def test_white_furnace_closed_form_on_the_gpu():
    from furnace_util import check_closed_form
    check_closed_form(lambda n: CLContext(n), width=96, height=72, iterations=256)


def test_gpu_agrees_with_the_libm_math_oracle_within_monte_carlo_error():
    """The CUDA path shares include/flx_math.h with the oracle it is bit-compared with; the "lm_" oracle does not (glibc's sinf,
    cosf, acosf, atan2f, powf).  Converged images of the two must agree -- this is the check that could see a bias in the header."""
    from furnace_util import render
    from oracle.oracle_host import RefContext, ref_available
    from test_oracle_cpu import open_room
    if not ref_available():
        pytest.skip("oracle/_ref not built")
    scene = open_room()
    W, H, N = 64, 48, 64 * 48
    rng = np.random.default_rng(5)
    env = EnvMapData.from_rgb(rng.uniform(0.1, 0.5, size=(16, 32, 3)).astype(np.float32))
    its = 1200
    params = room_params(scene, W, H, max_bounces=5, separate_queues=True, use_env_map=True, use_area_light=True, env_map_strength=1.5)
    with CLContext(N) as gpu:
        a, ca = render(gpu, scene, params, env, its)
    b, cb = render(RefContext(N, variant="lm_", parallel=True), scene, params, env, its)
    a, b = a.astype(np.float64), b.astype(np.float64)
    assert abs(ca.sum() - cb.sum()) < 0.01 * ca.sum()
    mean_rel = np.abs(a.mean(axis=0) - b.mean(axis=0)) / a.mean(axis=0)
    assert (mean_rel < 0.02).all(), "image means differ by %r" % (mean_rel,)

    def blocks(x):
        return x.reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
    blk = np.abs(blocks(a) - blocks(b)) / np.maximum(blocks(a), 1e-3)
    assert np.median(blk) < 0.03 and blk.max() < 0.5, (float(np.median(blk)), float(blk.max()))
