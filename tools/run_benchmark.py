#!/usr/bin/env python
"""The reference's own benchmark, Tracer::runBenchmark (src/tracer.cpp:362-528), on this library: 1024 x 1024, the
reference's three scenes (egyptcat, conference, country_kitchen), its per-iteration loop with a queue synchronisation every
iteration, Mrays/s logged every 0.5 s into the CSV format `scene;time;primary;extension;shadow;total;samples` that the
reference's plot_benchmarks.py reads, plus its one-line summary per scene.  The reference renders 30 s per scene; the default
here is 3 s (--seconds).  --integrator mk runs the microkernel branch of the same loop.  Cameras: the reference uses whatever
state file exists for a scene and otherwise its default camera; no state files are shipped, so the cameras of
tests/bench_configs.py are used (default camera for egyptcat)."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--integrator", choices=("wf", "mk"), default="wf")
    ap.add_argument("--scenes", default="egyptcat,conference,country_kitchen")
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    from bench_configs import ENV_MAPS, CONFIGS, teapot_params
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    W = H = 1024
    rows = ["scene;time;primary;extension;shadow;total;samples"]
    scenes_dir = os.path.join(ROOT, "oracle", "_ref", "scenes")
    for name in a.scenes.split(","):
        scene = SceneData.load_blob(os.path.join(scenes_dir, name + ".bin"))
        params = (CONFIGS[name] if name in CONFIGS else teapot_params)(scene, W, H)  # teapot_params = the reference's default camera and light
        if name not in CONFIGS:
            params.maxBounces = 8
        with CLContext(1 << 20) as ctx:  # wfBufferSize default (src/settings.cpp:20)
            ctx.uploadSceneData(scene)
            if name in ENV_MAPS:
                ctx.createEnvMap(EnvMapData.load_blob(os.path.join(scenes_dir, ENV_MAPS[name] + ".env.bin")))
            ctx.setupPixelStorage(W, H)
            tr = Tracer(ctx, params)
            r, s = tr.runBenchmarkScene("assets/%s" % name, a.seconds, use_wavefront=(a.integrator == "wf"))
            rows += r
            print("%s: %.1fM primary, %.2fM extension, %.2fM shadow, %.2fM samples, total: %.2fM rays/s  (%d iterations in %.2f s)" % (
                name, s["primary"], s["extension"], s["shadow"], s["samples"], s["total"], s["iterations"], s["seconds"]), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            f.write("\n".join(rows) + "\n")


if __name__ == "__main__":
    main()
