#!/bin/bash
mkdir -p gpurun_out
cd tests
timeout 1200 python - <<'PY' 2>&1 | tail -5
import sys, time
sys.path.insert(0, '..'); sys.path.insert(0, '.')
from fluctus_b200 import CLContext, SceneData
from conftest import scene_blob
from parity_util import run_lockstep
from oracle.oracle_host import RefContext
from bench_configs import conference_params, kitchen_params, luxball_params
t = time.time()
scene = SceneData.load_blob(scene_blob("conference"))
W, H, N = 1920, 1080, 1 << 21
with CLContext(N) as gpu:
    tg, tc = run_lockstep(gpu, RefContext(N, parallel_trace=True), scene, conference_params(scene, W, H), iterations=12, check_every=4)
    print("metric size, 12 iterations in lockstep with the oracle: OK", tg.stats, "%.0f s" % (time.time() - t))
t = time.time()
scene = SceneData.load_blob(scene_blob("luxball"))
N = 1 << 19
with CLContext(N) as gpu:
    tg, tc = run_lockstep(gpu, RefContext(N, parallel_trace=True), scene, luxball_params(scene, 1280, 720), iterations=24, check_every=8)
    print("luxball 1280x720, 2^19 paths, 24 iterations in lockstep: OK", tg.stats, "%.0f s" % (time.time() - t))
PY
