"""CPU tests of the host-side pieces: pinned math, layouts, the C ABI surface, scene containers, the loop driver."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import fluctus_b200 as fx
from fluctus_b200 import _lib
from fluctus_b200.scene import build_bvh, make_room_scene, room_params
from oracle.oracle_host import PORT_LIB, REF_LIB, port_available, ref_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_port = pytest.mark.skipif(not port_available(), reason="oracle/liboracle.so not built")


def _ulp_err(got, ref64):
    sp = np.spacing(np.abs(ref64.astype(np.float32))).astype(np.float64)
    return np.nanmax(np.abs(got.astype(np.float64) - ref64) / sp)


@needs_port
def test_pinned_math_within_one_ulp_of_float64():
    """include/flx_math.h (the functions every consumer shares) vs numpy float64; tolerance 1 ulp, observed <= 0.5."""
    lib = C.CDLL(PORT_LIB)
    lib.port_math.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
    rng = np.random.default_rng(0)

    def run(fn, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = np.ascontiguousarray(b if b is not None else a, np.float32)
        o = np.empty_like(a)
        lib.port_math(fn, a.ctypes.data, b.ctypes.data, o.ctypes.data, len(a))
        return o

    x = np.concatenate([rng.uniform(-10, 10, 400000), rng.uniform(-2000, 2000, 50000), [0, np.pi, np.pi / 2, 2 * np.pi]]).astype(np.float32)
    assert _ulp_err(run(0, x), np.sin(x.astype(np.float64))) <= 1.0
    assert _ulp_err(run(1, x), np.cos(x.astype(np.float64))) <= 1.0
    xt = rng.uniform(-1.5, 1.5, 200000).astype(np.float32)
    assert _ulp_err(run(2, xt), np.tan(xt.astype(np.float64))) <= 1.0
    xa = np.concatenate([rng.uniform(-1, 1, 200000), [1, -1, 0]]).astype(np.float32)
    assert _ulp_err(run(3, xa), np.arccos(xa.astype(np.float64))) <= 1.0
    y, xx = rng.normal(size=200000).astype(np.float32), rng.normal(size=200000).astype(np.float32)
    assert _ulp_err(run(4, y, xx), np.arctan2(y.astype(np.float64), xx.astype(np.float64))) <= 1.0
    xp = np.concatenate([rng.uniform(0, 1, 200000), rng.uniform(0, 100, 5000), [0, 1]]).astype(np.float32)
    e = np.full_like(xp, 2.2)
    assert _ulp_err(run(5, xp, e), np.power(xp.astype(np.float64), np.float64(np.float32(2.2)))) <= 1.0
    # special values the path relies on
    assert run(5, [0.0], [2.2])[0] == 0.0 and run(5, [1.0], [2.2])[0] == 1.0
    assert run(4, [0.0], [-1.0])[0] == np.float32(np.pi) and run(4, [0.0], [0.0])[0] == 0.0
    assert np.isnan(run(3, [1.5])[0])


@needs_port
def test_rng_known_answers():
    """hash/rand of src/random.cl:7-22: first outputs from seed 0, computed independently in Python integers."""
    def h(s):
        s = ((s ^ 61) ^ (s >> 16)) & 0xffffffff
        s = (s * 9) & 0xffffffff
        s = s ^ (s >> 4)
        s = (s * 0x27d4eb2d) & 0xffffffff
        return s ^ (s >> 15)
    lib = C.CDLL(PORT_LIB)
    out, seeds = np.empty(8, np.float32), np.empty(8, np.uint32)
    lib.port_rand(C.c_uint32(0), out.ctypes.data_as(C.c_void_p), seeds.ctypes.data_as(C.c_void_p), 8)
    s, exp = 0, []
    for _ in range(8):
        s = h(s)
        exp.append(s)
    assert list(seeds) == exp
    assert np.array_equal(out, (np.array(exp, np.uint32).astype(np.float32) * np.float32(1.0 / 4294967296.0)))
    assert (out >= 0).all() and (out <= 1).all()


def test_struct_layouts_match_reference_sizes():
    assert C.sizeof(fx.RenderParams) == 240 and fx.RenderParams.camera.offset == 96 and fx.RenderParams.width.offset == 184
    assert fx.RenderParams.worldRadius.offset == 228 and fx.RenderParams.maxBounces.offset == 208
    assert fx.NODE_DTYPE.itemsize == 48 and fx.NODE_DTYPE.fields["nPrims"][1] == 40 and fx.NODE_DTYPE.fields["link"][1] == 36
    assert fx.TRIANGLE_DTYPE.itemsize == 160 and fx.TRIANGLE_DTYPE.fields["matId"][1] == 144
    assert fx.MATERIAL_DTYPE.itemsize == 80 and fx.MATERIAL_DTYPE.fields["type"][1] == 68 and fx.MATERIAL_DTYPE.fields["Ns"][1] == 48
    if ref_available():  # sizeof() as the reference's own geom.h compiles
        lib = C.CDLL(REF_LIB)
        out = (C.c_uint32 * 8)()
        lib.ref_layout(out)
        assert list(out) == [256, 48, 160, 80, 240, 32, 12, 64]


def test_c_abi_library_exports_every_declared_symbol():
    """Every function declared in include/fluctus_b200.h is exported by the built library and bound in _lib.py."""
    header = open(os.path.join(ROOT, "include", "fluctus_b200.h")).read()
    declared = set(re.findall(r"\b(flx_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert os.path.exists(_lib.LIB_PATH), "libfluctus_b200.so not built (python fluctus_b200/csrc/build.py)"
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (flx_[a-z0-9_]+)", nm))
    assert declared <= exported, "declared but not exported: %s" % sorted(declared - exported)
    assert declared == set(_lib.EXPORTS), "header and ctypes binding disagree: %s" % sorted(declared ^ set(_lib.EXPORTS))
    lib = _lib.load()  # loads without a GPU; no compute calls here
    assert b"sm_100a" in lib.flx_version()


def test_library_has_only_sm100a_code_and_no_oracle_symbols():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out and not re.search(r"sm_(?!100a)\d+", out), out
    nm = subprocess.run(["nm", "-D", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "port_" not in nm and "ref_" not in nm.replace("deref", "")


def test_create_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(fx.FluctusError, match="no CUDA device|no CPU path"):
        fx.CLContext(1024)


def test_room_scene_and_bvh_builder_are_well_formed():
    s = make_room_scene(materials="mixed", textured=True)
    n = s.nodes
    inner = n["nPrims"] == 0
    assert inner[0] and n["parent"][0] == -1
    assert (n["link"][inner] > np.flatnonzero(inner) + 1).all() and (n["link"][inner] < len(n)).all()
    leaves = ~inner
    assert n["nPrims"][leaves].sum() == len(s.indices) and sorted(s.indices) == list(range(len(s.tris)))
    # children lie inside parents
    for i in np.flatnonzero(inner)[:200]:
        for c in (i + 1, n["link"][i]):
            assert (n["bmin"][c][:3] >= n["bmin"][i][:3]).all() and (n["bmax"][c][:3] <= n["bmax"][i][:3]).all()
    assert s.material_types == 0x7e
    p = room_params(s, 32, 24)
    assert p.width == 32 and p.useAreaLight == 1 and abs(p.worldRadius - s.world_radius) < 1e-7


def test_env_tables_are_a_valid_alias_method():
    rng = np.random.default_rng(1)
    rgb = rng.uniform(0, 1, size=(8, 16, 3)).astype(np.float32)
    env = fx.EnvMapData.from_rgb(rgb)
    n = 8 * 16
    assert abs(env.pdf.mean() - 1.0) < 1e-4
    # alias method reconstructs the pdf: P(i) = (prob[i] + sum_{j: alias[j]=i} (1-prob[j])) / n
    recon = env.prob.astype(np.float64).copy()
    for j in range(n):
        if env.prob[j] < 1.0:
            recon[env.alias[j]] += 1.0 - env.prob[j]
    assert np.allclose(recon, env.pdf, atol=2e-4)


def test_tracer_replays_reference_call_order():
    """Tracer.start/iterate issue CLContext calls in the order of src/tracer.cpp:236-240 and 433-439/447/465."""
    calls = []

    class Fake:
        def __getattr__(self, name):
            def f(*a, **k):
                calls.append(name)
            return f

        def tilePixels(self):
            return 12

    t = fx.Tracer(Fake(), fx.RenderParams())
    t.start()
    assert calls == ["updateParams", "resetPixelIndex", "enqueueWfResetKernel", "enqueueWfRaygenKernel", "enqueueWfExtRayKernel", "enqueueClearWfQueues", "finishQueue"]
    calls.clear()
    t.iterate()
    assert calls == ["enqueueWfLogicKernel", "enqueueWfRaygenKernel", "enqueueWfMaterialKernels", "enqueueGetCounters", "enqueueWfExtRayKernel",
                     "enqueueWfShadowRayKernel", "enqueueClearWfQueues", "enqueuePostprocessKernel", "finishQueue", "updatePixelIndex"]


def test_env_tables_reproduce_reference_builder_bit_for_bit():
    """EnvMapData.from_rgb (mirror of EnvironmentMap::computeProbabilities, src/envmap.cpp:31-114) against the tables the
    reference's own envmap.cpp produced for assets/env_maps/night.hdr (oracle/_ref/scenes/night.env.bin)."""
    from conftest import SCENES_DIR
    p = os.path.join(SCENES_DIR, "night.env.bin")
    if not os.path.exists(p):
        pytest.skip("env map blob not built (needs /root/reference)")
    ref = fx.EnvMapData.load_blob(p)
    mine = fx.EnvMapData.from_rgb(ref.rgb)
    assert np.array_equal(mine.pdf, ref.pdf) and np.array_equal(mine.prob, ref.prob) and np.array_equal(mine.alias, ref.alias)


@pytest.mark.skipif(not port_available(), reason="oracle/liboracle.so not built")
@pytest.mark.parametrize("wavefront", [True, False])
def test_benchmark_protocol_replays_the_reference_loop_and_csv(wavefront):
    """Tracer.runBenchmarkScene = one scene of Tracer::runBenchmark (src/tracer.cpp:372-383, 416-503), here driven on the CPU
    oracle with a fake clock: the CSV rows have the reference's seven columns (what plot_benchmarks.py parses), the logged
    intervals add up to the totals, both integrators run, and the wavefront start-up without a prologue works (every path is
    found at length 0 by the first logic pass and regenerated)."""
    from oracle.oracle_host import PortContext
    scene = make_room_scene(materials="diffuse")
    W, H, N = 24, 16, 24 * 16
    params = room_params(scene, W, H, max_bounces=2)
    ctx = PortContext(N)
    ctx.uploadSceneData(scene)
    ctx.setupPixelStorage(W, H)
    ticks = iter(np.arange(0.0, 100.0, 0.2))
    tr = fx.Tracer(ctx, params)
    rows, summary = tr.runBenchmarkScene("assets/room", render_len=2.0, use_wavefront=wavefront, log_every=0.5, clock=lambda: float(next(ticks)))
    assert len(rows) >= 2 and all(len(r.split(";")) == 7 and r.startswith("assets/room;") for r in rows)
    vals = np.array([[float(x) for x in r.split(";")[1:]] for r in rows])
    assert np.allclose(vals[:, 4], vals[:, 1] + vals[:, 2] + vals[:, 3])  # total = primary + extension + shadow
    assert summary["iterations"] >= 5 and summary["primary"] > 0 and summary["extension"] > 0 and summary["shadow"] > 0 and summary["samples"] > 0
    assert abs(summary["total"] - (summary["primary"] + summary["extension"] + summary["shadow"])) < 1e-9
    pix = ctx.readPixels()
    assert pix[:, 3].sum() > 0 and np.isfinite(pix).all()
