#!/usr/bin/env python
"""A short steady-state run of the metric workload (conference 1920x1080, N = 2^21) for ncu: warm-up iterations, then a few
iterations to capture.  Usage on the GPU box (B200_PROFILING.md):

  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python tools/prof_step.py
  ncu --set full --clock-control none --import-source on -k regex:'k_trace_persistent|k_logic' -s 24 -c 3 -o gpurun_out/prof python tools/prof_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scene", default="conference")
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--tasks", type=int, default=1 << 21)
    ap.add_argument("--warmup", type=int, default=12)
    ap.add_argument("--iters", type=int, default=2)
    ap.add_argument("--tune", default="")
    ap.add_argument("--builder", default="", help="fast | ploc: trace through a GPU-built hierarchy instead of the scene file's SBVH")
    a = ap.parse_args()
    from bench_configs import ENV_MAPS, params_for
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    scene = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", a.scene + ".bin"))
    params = params_for(a.scene, scene, a.width, a.height)
    with CLContext(a.tasks) as ctx:
        if a.tune:
            ctx.setTuning(**{k: int(v) for k, v in (kv.split("=") for kv in a.tune.split(","))})
        if a.builder:
            nodes, idx, _ = ctx.buildBVH(scene.tris, 8, a.builder)
            scene = SceneData(scene.tris, idx, nodes, scene.materials, scene.tex_desc, scene.tex_data)
        ctx.uploadSceneData(scene)
        if a.scene in ENV_MAPS:
            ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", ENV_MAPS[a.scene] + ".env.bin")))
        ctx.setupPixelStorage(a.width, a.height)
        tr = Tracer(ctx, params)
        tr.start()
        ctx.render(a.warmup)
        ctx.finishQueue()
        ms = ctx.renderTimed(a.iters)
        st = ctx.getStats()
        print("prof_step: %d iterations, %.3f ms each" % (a.iters, ms / a.iters))


if __name__ == "__main__":
    main()
