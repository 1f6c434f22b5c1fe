"""The white-furnace scene of the self-consistency pins: ONE convex object with a Lambertian albedo under a uniform
environment -- the case with a closed form.

A ray leaving a convex object never meets it again, so every camera path is "camera -> object -> environment" (or straight to
the environment), and with radiance L everywhere in the environment the rendering equation collapses to

    pixel = L                       where the pixel sees the environment,
    pixel = albedo * L              where it sees the object            (albedo = Kd ^ 2.2: matGetAlbedo, src/utils.cl:136-141),

for EVERY unbiased estimator: BSDF sampling only (sampleImpl), light sampling only (sampleExpl: alias-method env-map samples),
and both with the balance heuristic (MIS).  With cosine-weighted BSDF sampling f * cos / pdf is exactly the albedo, so the
implicit-only estimator has zero variance: every sample of an object pixel is albedo * L up to rounding.  The other two converge
to the same numbers -- up to the quadrature error of the reference's light sampler, which only ever returns texel centres (found
by this test: +2.7 % on a 16 x 8 map, < 1 % on the 128 x 64 map used here)."""
import numpy as np

from fluctus_b200 import EnvMapData, SceneData, look_at, make_params
from fluctus_b200.scene import _material, _tri, build_bvh
from fluctus_b200.structs import MATERIAL_DTYPE, TRIANGLE_DTYPE

KD = 0.5
ALBEDO = float(np.float32(KD) ** np.float32(2.2))
ENV = (0.7, 0.9, 1.1)
STRENGTH = 2.0


def cube_scene():
    c = [(-1, -1, -1), (1, -1, -1), (1, 1, -1), (-1, 1, -1), (-1, -1, 1), (1, -1, 1), (1, 1, 1), (-1, 1, 1)]
    quads = [(0, 3, 2, 1), (4, 5, 6, 7), (0, 1, 5, 4), (2, 3, 7, 6), (1, 2, 6, 5), (0, 4, 7, 3)]  # outward-facing
    tris = []
    for a, b, cc, d in quads:
        tris.append(_tri(c[a], c[b], c[cc], 0))
        tris.append(_tri(c[a], c[cc], c[d], 0))
    tris = np.array(tris, TRIANGLE_DTYPE)
    nodes, indices = build_bvh(tris, max_leaf=2)
    return SceneData(tris, indices, nodes, np.array([_material(kd=(KD, KD, KD))], MATERIAL_DTYPE))


def uniform_env(w=128, h=64):
    # fine enough that the reference's light sampling -- which always returns the CENTRE of the texel it picked
    # (sampleEnvMapAlias, src/env_map.cl:65-90), i.e. integrates by a midpoint rule over w x h directions -- is within a few 1e-3
    rgb = np.empty((h, w, 3), np.float32)
    rgb[...] = ENV
    return EnvMapData.from_rgb(rgb)


def furnace_params(scene, width, height, sample_impl, sample_expl, max_bounces=3):
    cam = look_at((3.1, 2.3, 4.2), (0.0, 0.0, 0.0), fov=40.0)
    return make_params(width, height, cam, scene.world_radius, len(scene.tris), light=False, max_bounces=max_bounces, use_env_map=True,
                       env_map_strength=STRENGTH, sample_impl=sample_impl, sample_expl=sample_expl)


def render(ctx, scene, params, env, iterations):
    from fluctus_b200 import Tracer
    ctx.uploadSceneData(scene)
    ctx.createEnvMap(env)
    ctx.setupPixelStorage(params.width, params.height)
    tr = Tracer(ctx, params)
    tr.start()
    for _ in range(iterations):
        tr.iterate()
    pix = ctx.readPixels()
    assert (pix[:, 3] > 0).all()
    return pix[:, :3] / pix[:, 3:4], pix[:, 3]


def check_closed_form(make_ctx, width=40, height=30, iterations=160):
    """make_ctx(n) -> a context with CLContext's method set.  Renders the furnace with the three estimators and checks them
    against the closed form and against each other.  Returns the three mean images."""
    scene, env = cube_scene(), uniform_env()
    L = np.asarray(ENV, np.float64) * STRENGTH
    n = width * height
    out = {}
    for name, (impl, expl) in (("implicit", (True, False)), ("explicit", (False, True)), ("mis", (True, True))):
        ctx = make_ctx(n)
        out[name], _ = render(ctx, scene, furnace_params(scene, width, height, impl, expl), env, iterations)
        if hasattr(ctx, "close"):
            ctx.close()
    imp = out["implicit"].astype(np.float64)
    # zero-variance estimator: every pixel is a mix f * L + (1 - f) * albedo * L of the two closed-form values, per channel
    f = (imp / L - ALBEDO) / (1.0 - ALBEDO)
    assert (f > -1e-4).all() and (f < 1 + 1e-4).all(), "a pixel lies outside [albedo * L, L]: %r .. %r" % (f.min(), f.max())
    assert np.abs(f - f[:, :1]).max() < 1e-4, "the three colour channels disagree on the coverage of a pixel"
    # "inside" / "outside" = the pixel AND its eight neighbours see only the object / only the environment: the three renders
    # jitter their samples differently (path lengths differ, so the pixel counter advances differently), and a pixel next to the
    # silhouette may catch a sample of the other kind in one render and not in another
    def interior(mask):
        m = mask.reshape(height, width)
        core = m.copy()
        for dy in (-1, 0, 1):
            for dx in (-1, 0, 1):
                core &= np.roll(np.roll(m, dy, axis=0), dx, axis=1)
        core[0, :] = core[-1, :] = core[:, 0] = core[:, -1] = False
        return core.reshape(-1)
    inside, outside = interior(f[:, 0] < 1e-4), interior(f[:, 0] > 1 - 1e-4)
    assert inside.sum() > n // 10 and outside.sum() > n // 10, "camera does not see enough of both object and environment"
    assert np.abs(imp[inside] / (ALBEDO * L) - 1).max() < 2e-5, "object pixels are not albedo * L"
    assert np.abs(imp[outside] / L - 1).max() < 2e-5, "environment pixels are not L"
    # the other two estimators converge to the same picture: globally, and per pixel within their Monte-Carlo error
    for name in ("explicit", "mis"):
        img = out[name].astype(np.float64)
        assert np.abs(img[outside] / L - 1).max() < 2e-5, "%s: environment pixels are not L" % name
        rel = img[inside] / (ALBEDO * L) - 1
        assert abs(rel.mean()) < 0.015, "%s sampling is biased on the object: mean relative error %.4f" % (name, rel.mean())
        rms = float(np.sqrt((rel ** 2).mean()))
        assert rms < 0.3 and np.abs(rel).max() < 1.5, "%s: pixels scatter too far around the closed form (rms %.3f, max %.3f)" % (name, rms, np.abs(rel).max())
    return out
