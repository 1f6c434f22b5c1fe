#!/usr/bin/env python
"""Throughput of the microkernel integrator (Tracer::renderSingle, src/tracer.cpp:95-169) at full size on one GPU.
One JSON line per configuration: Mrays/s = (primary + extension + shadow rays) / device time of the fused sample loop,
per-kernel CUDA-event times, and the same scene through the wavefront loop for comparison."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scenes", default="conference,luxball,country_kitchen")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--spp", type=int, default=16)
    a = ap.parse_args()
    import numpy as np
    from bench_configs import ENV_MAPS, params_for
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    W, H = a.width, a.height
    for scene_name in a.scenes.split(","):
        scene = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", scene_name + ".bin"))
        params = params_for(scene_name, scene, W, H)
        with CLContext(W * H) as ctx:
            ctx.uploadSceneData(scene)
            if scene_name in ENV_MAPS:
                ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", ENV_MAPS[scene_name] + ".env.bin")))
            ctx.setupPixelStorage(W, H)
            tr = Tracer(ctx, params)
            tr.renderSingle(2, fused=True)  # warm-up
            ctx.updateParams(params)
            ctx.enqueueResetKernel(params)
            ctx.finishQueue()
            ctx.resetStats()
            ctx.setProfiling(True)
            ctx.timerBegin()
            ctx.renderSingleLoop(a.spp)
            ms = ctx.timerEnd()
            ctx.setProfiling(False)
            st = ctx.getStats()
            perf = ctx.checkTracingPerf()
            pix = ctx.readPixels()
            rays = st.primaryRays + st.extensionRays + st.shadowRays
            row = dict(integrator="microkernel", scene=scene_name, width=W, height=H, spp=a.spp, max_bounces=params.maxBounces,
                       mrays_per_s=round(rays / ms / 1e3, 1), ms_per_sample=round(ms / a.spp, 3), msamples_per_s=round(st.samples / ms / 1e3, 1),
                       primary=int(st.primaryRays), extension=int(st.extensionRays), shadow=int(st.shadowRays),
                       kernel_ms_per_sample={k: round(v[0] / a.spp, 4) for k, v in perf.items() if v[1]},
                       exact_spp=bool(np.array_equal(pix[:, 3], np.full(W * H, float(a.spp), np.float32))),
                       nonfinite_pixels=int((~np.isfinite(pix)).any(axis=1).sum()))
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
