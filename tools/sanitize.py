#!/usr/bin/env python
"""Small end-to-end run for compute-sanitizer (memcheck / racecheck / initcheck): every kernel, both queue modes, env map,
tiling with one part.  Usage on the GPU box:  compute-sanitizer --tool memcheck python tools/sanitize.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from fluctus_b200 import CLContext, EnvMapData, Tracer  # noqa: E402
from fluctus_b200.scene import make_room_scene, room_params  # noqa: E402


def main():
    scene = make_room_scene(materials="mixed", textured=True)
    rng = np.random.default_rng(3)
    env = EnvMapData.from_rgb(rng.uniform(0.0, 0.4, size=(16, 32, 3)).astype(np.float32))
    for variant in (0, 1, 2, 3):
        for separate in (False, True):
            params = room_params(scene, 40, 24, max_bounces=4, separate_queues=separate, use_env_map=True, env_map_strength=1.5)
            with CLContext(1500) as ctx:
                ctx.setTuning(trace_variant=variant)
                ctx.uploadSceneData(scene)
                ctx.createEnvMap(env)
                ctx.setupPixelStorage(40, 24)
                ctx.setTile(0, 1, 4)
                tr = Tracer(ctx, params)
                tr.start()
                for _ in range(3):
                    tr.iterate()
                ctx.render(3)
                ctx.setCounting(True)
                ctx.render(1)
                ctx.setCounting(False)
                tr.update()
                pix = ctx.readPixels()
                assert np.isfinite(pix).all()
                # the microkernel integrator on the same context: renderSingle call by call and fused, then the interactive loop
                tr.renderSingle(2)
                tr.renderSingle(2, fused=True)
                tr.iteration = 0
                tr.updateMicrokernel()
                tr.updateMicrokernel()
                pix = ctx.readPixels()
                assert np.isfinite(pix).all()
    # the hierarchy builder, incl. the one- and two-triangle cases and a scene rendered through its tree
    with CLContext(1500) as ctx:
        for quality in ("fast", "ploc", "ploc_opt"):
            for n in (1, 2, len(scene.tris)):
                nodes, idx, _ = ctx.buildBVH(scene.tris[:n], 8 if n > 2 else 1, quality)
        from fluctus_b200 import SceneData
        built = SceneData(scene.tris, idx, nodes, scene.materials, scene.tex_desc, scene.tex_data)
        params = room_params(built, 40, 24, max_bounces=3)
        ctx.uploadSceneData(built)
        ctx.setupPixelStorage(40, 24)
        tr = Tracer(ctx, params)
        tr.start()
        ctx.render(3)
        ctx.setTuning(fuse_stages=0, overlap_postprocess=2, repack_on_host=1)
        ctx.uploadSceneData(built)
        tr.start()
        ctx.render(2)
        tr.iterate()
        assert np.isfinite(ctx.readPixels()).all()
    # round 2: denoiser feature buffers in both integrators, the gather on its own stream (one-rank communicator) in both
    # placements, a re-upload into kept allocations, checkpoint save / load, page-locked host buffers, thin lens, 128-path logic tiles
    import tempfile
    from fluctus_b200 import FluctusError, look_at, make_params, pinned_empty
    with CLContext(1500) as ctx:
        ctx.setDenoiser(True)
        params = room_params(scene, 40, 24, max_bounces=4, separate_queues=True)
        for rep in range(2):  # the second round re-uploads into the allocations of the first
            ctx.uploadSceneData(scene.pinned())
            ctx.setupPixelStorage(40, 24)
            tr = Tracer(ctx, params)
            tr.start()
            for _ in range(3):
                tr.iterate()
            ctx.render(2)
            out = pinned_empty((40 * 24, 4), np.float32)
            ctx.readPixels(out)
            assert np.isfinite(out).all() and np.isfinite(ctx.readDenoiserAOV("normal", True)).all() and np.isfinite(ctx.readDenoiserAOV("albedo")).all()
            tr.renderSingle(2)
        ck = os.path.join(tempfile.mkdtemp(), "s.ckpt")
        tr = Tracer(ctx, params)
        tr.start()
        ctx.render(2)
        ctx.saveCheckpoint(ck)
        ctx.render(1)
        ctx.loadCheckpoint(ck)
        ctx.render(1)
        ctx.setTuning(logic_tile=128)
        cam = look_at((0.0, 1.0, 0.95), (0.0, 0.9, -0.2), fov=70.0, aperture=0.02, focal_dist=1.2)
        p2 = make_params(40, 24, cam, scene.world_radius, n_tris=len(scene.tris), max_bounces=3)
        tr = Tracer(ctx, p2)
        tr.start()
        ctx.render(3)
    for direct in (0, 1):
        with CLContext(1500) as ctx:
            try:
                ctx.setTuning(gather_direct=direct)
                ctx.setTile(0, 1, 4)
                ctx.commInit(ctx.commUniqueId(), 0, 1)
            except FluctusError as e:
                print("gather not exercised: %s" % e)
                break
            ctx.uploadSceneData(scene)
            for (w, h) in ((40, 22), (40, 24)):
                ctx.setupPixelStorage(w, h)
                tr = Tracer(ctx, room_params(scene, w, h, max_bounces=3))
                tr.start()
                full = np.zeros((w * h, 4), np.float32)
                for it in range(4):
                    ctx.render(1)
                    ctx.gatherPixels(0, full if it % 2 else None)
                ctx.finishQueue()
                assert np.isfinite(full).all()
    print("SANITIZE_RUN_OK")


if __name__ == "__main__":
    main()
