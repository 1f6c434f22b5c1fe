"""Where the end-to-end time of bench.py's e2e leg goes (GPU box): context creation, scene upload (+ hierarchy repack), image
allocation, prologue, the per-stage loop of Tracer.iterate() and the image read-back, next to the fused loop's per-iteration time.
FLX_DEBUG_TIMING=1 additionally prints flx_upload_scene's own split."""
import os, sys, time
ROOT = os.getcwd()
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from bench_configs import conference_params
from fluctus_b200 import CLContext, SceneData, Tracer
scene = SceneData.load_blob(os.path.join(ROOT, "oracle/_ref/scenes/conference.bin"))
params = conference_params(scene, 1920, 1080)
for rep in range(2):
    t = [time.perf_counter()]
    ctx = CLContext(1 << 21); t.append(time.perf_counter())
    ctx.uploadSceneData(scene); t.append(time.perf_counter())
    ctx.setupPixelStorage(1920, 1080); t.append(time.perf_counter())
    tr = Tracer(ctx, params); tr.start(); t.append(time.perf_counter())
    for _ in range(20): tr.iterate()
    t.append(time.perf_counter())
    for _ in range(500): tr.iterate()
    t.append(time.perf_counter())
    img = ctx.readPixels(); t.append(time.perf_counter())
    names = ["create", "uploadSceneData", "setupPixelStorage", "start", "20 iterations", "500 iterations", "readPixels"]
    print(rep, {n: round((b - a) * 1e3, 2) for n, a, b in zip(names, t, t[1:])}, "ms; per-iteration %.4f ms" % ((t[6] - t[5]) / 500 * 1e3))
    ms = ctx.renderTimed(200)
    print("   fused loop per iteration %.4f ms" % (ms / 200))
    ctx.close()
