#!/usr/bin/env python
"""Experiment (GPU box): how much faster do the traversal kernels run when the ray queues are ordered for coherence?
The queue ORDER never changes results (only which lane traces which ray), so this measures the head-room of a ray
re-ordering pass before one is written.  Orders the extension / shadow queue on the host by several keys and times the
kernel on each.  One JSON line per (queue, key)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from bench_configs import params_for  # noqa: E402
from fluctus_b200 import CLContext, SLOT, SceneData, Tracer  # noqa: E402
from fluctus_b200.structs import QueueCounters  # noqa: E402


def morton3(x, y, z, bits):
    def spread(v):
        v = v.astype(np.uint64)
        out = np.zeros_like(v)
        for b in range(bits):
            out |= ((v >> np.uint64(b)) & np.uint64(1)) << np.uint64(3 * b)
        return out
    return spread(x) | (spread(y) << np.uint64(1)) | (spread(z) << np.uint64(2))


def keys(tasks, q, o_slot, d_slot, lo, hi):
    o = tasks[o_slot:o_slot + 3][:, q].view(np.float32)
    d = tasks[d_slot:d_slot + 3][:, q].view(np.float32)
    octant = ((d[0] < 0).astype(np.uint64) | ((d[1] < 0).astype(np.uint64) << np.uint64(1)) | ((d[2] < 0).astype(np.uint64) << np.uint64(2)))
    out = {"as_pushed": None, "path_index": q.astype(np.uint64), "octant": octant}
    for bits in (2, 3, 4, 6, 10):
        n = 1 << bits
        c = [np.clip(((o[k] - lo[k]) / max(hi[k] - lo[k], 1e-20) * n).astype(np.int64), 0, n - 1) for k in range(3)]
        m = morton3(c[0], c[1], c[2], bits)
        out["cell%d_octant" % n] = (m << np.uint64(3)) | octant
        out["octant_cell%d" % n] = (octant << np.uint64(3 * bits)) | m
    # direction-major: 6-bit direction cell (cube of 4^3 over [-1,1]^3) then origin cell 8^3
    dc = [np.clip(((d[k] + 1.0) * 0.5 * 4).astype(np.int64), 0, 3) for k in range(3)]
    c = [np.clip(((o[k] - lo[k]) / max(hi[k] - lo[k], 1e-20) * 8).astype(np.int64), 0, 7) for k in range(3)]
    out["cell8_dir64"] = (morton3(c[0], c[1], c[2], 3) << np.uint64(6)) | morton3(dc[0], dc[1], dc[2], 2)
    return out


def main():
    scene_name = sys.argv[1] if len(sys.argv) > 1 else "conference"
    W, H, N = 1920, 1080, 1 << 21
    scene = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", scene_name + ".bin"))
    params = params_for(scene_name, scene, W, H)
    lo = [float(scene.nodes["bmin"][0][k]) for k in range(3)]
    hi = [float(scene.nodes["bmax"][0][k]) for k in range(3)]
    with CLContext(N) as ctx:
        ctx.uploadSceneData(scene)
        if scene_name == "country_kitchen":
            from fluctus_b200 import EnvMapData
            ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", "night.env.bin")))
        ctx.setupPixelStorage(W, H)
        tr = Tracer(ctx, params)
        tr.start()
        ctx.render(40)
        ctx.finishQueue()
        # one more iteration up to the point where the queues are full
        cnt = QueueCounters()
        ctx.enqueueWfLogicKernel(params, False)
        ctx.enqueueWfRaygenKernel(params)
        ctx.enqueueWfMaterialKernels(params)
        ctx.enqueueGetCounters(cnt)
        ctx.finishQueue()
        tasks = ctx.readTasks()
        for qname, n, o_slot, d_slot, launch, kid in (("extension", cnt.extensionQueue, SLOT.ORIG, SLOT.DIR, ctx.enqueueWfExtRayKernel, "extrays"),
                                                      ("shadow", cnt.shadowQueue, SLOT.SHADOW_ORIG, SLOT.SHADOW_DIR, ctx.enqueueWfShadowRayKernel, "shadowrays")):
            q0 = ctx.readQueue(qname, n)
            for kname, key in keys(tasks, q0, o_slot, d_slot, lo, hi).items():
                q = q0 if key is None else q0[np.argsort(key, kind="stable")]
                ctx.writeQueue(qname, q)
                launch(params)  # warm
                ctx.finishQueue()
                ctx.resetStats()
                ctx.setProfiling(True)
                for _ in range(5):
                    launch(params)
                ctx.finishQueue()
                ctx.setProfiling(False)
                ms = ctx.checkTracingPerf()[kid][0] / 5
                print(json.dumps(dict(scene=scene_name, queue=qname, rays=int(n), order=kname, ms=round(ms, 4), mrays_per_s=round(n / ms / 1e3, 1))), flush=True)
            ctx.writeQueue(qname, q0)


if __name__ == "__main__":
    main()
