#!/usr/bin/env python
"""Where a GPU-built PLOC tree loses against the reference's SBVH: for either tree on the same scene, whole-loop throughput,
per-kernel device times (profiling mode: kernels serialised) and the traversal work counted on the benchmark's own rays
(node visits V, box tests B, triangle tests T, hit updates U per ray).  One JSON line per scene."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from bench_configs import ENV_MAPS, params_for
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    W, H, N = 1920, 1080, 1 << 21
    for scene_name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["conference"]):
        ref = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", scene_name + ".bin"))
        row = dict(scene=scene_name)
        with CLContext(N) as ctx:
            built = {}
            for q in ("ploc", "ploc_opt"):
                nodes, idx, _ = ctx.buildBVH(ref.tris, 8, q)
                built[q] = SceneData(ref.tris, idx, nodes, ref.materials, ref.tex_desc, ref.tex_data)
            for label, sc in (("sbvh", ref), ("ploc", built["ploc"]), ("ploc_opt", built["ploc_opt"]), ("sbvh_again", ref)):
                params = params_for(scene_name, sc, W, H)
                ctx.uploadSceneData(sc)
                if scene_name in ENV_MAPS:
                    ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", ENV_MAPS[scene_name] + ".env.bin")))
                ctx.setupPixelStorage(W, H)
                tr = Tracer(ctx, params)
                tr.start()
                ctx.render(20)
                ctx.finishQueue()
                ctx.resetStats()
                ms = ctx.renderTimed(100)
                st = ctx.getStats()
                out = dict(mrays_per_s=round((st.extensionRays + st.shadowRays) / ms / 1e3, 1), ms_per_iteration=round(ms / 100, 4))
                ctx.setTuning(overlap_trace=0)  # kernels one after the other: per-kernel times that add up
                ctx.resetStats()
                ctx.setProfiling(True)
                ctx.renderTimed(20)
                out["kernel_ms_serialised"] = {k: round(v[0] / 20, 4) for k, v in ctx.checkTracingPerf().items() if v[1]}
                ctx.setProfiling(False)
                ctx.setTuning(overlap_trace=1)
                ctx.setCounting(True)
                ctx.render(5)
                ctx.finishQueue()
                for kind, c in ctx.getTraceCounts().items():
                    r = max(c["rays"], 1)
                    out[kind + "_per_ray"] = dict(V=round(c["nodes"] / r, 2), B=round(c["boxes"] / r, 2), T=round(c["tris"] / r, 2), U=round(c["updates"] / r, 3))
                ctx.setCounting(False)
                row[label] = out
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
