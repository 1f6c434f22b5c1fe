#!/usr/bin/env python
"""GPU hierarchy build (flx_build_bvh) on the BASELINE scenes: device build time, tree statistics next to the reference's
SBVH (made by the reference's own builder, oracle/_ref/scenes), and wavefront throughput through either tree.
One JSON line per scene."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    from bench_configs import ENV_MAPS, params_for
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    from parity_util import validate_bvh
    W, H, N = 1920, 1080, 1 << 21
    for scene_name in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["conference", "luxball", "country_kitchen"]):
        ref = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", scene_name + ".bin"))
        row = dict(scene=scene_name, triangles=len(ref.tris))
        with CLContext(N) as ctx:
            built = {}
            for quality, cost in (("fast", 100), ("ploc", 100), ("ploc_opt", 100), ("ploc_opt", 150), ("ploc_opt", 200), ("ploc", 200)):
                ctx.setTuning(bvh_tri_cost=cost)
                ctx.buildBVH(ref.tris, 8, quality)  # warm-up (allocations, cub temp sizing)
                times = [ctx.buildBVH(ref.tris, 8, quality)[2] for _ in range(5)]
                nodes, idx, _ = ctx.buildBVH(ref.tris, 8, quality)
                key = quality if cost == 100 else "%s_tc%d" % (quality, cost)
                row["build_ms_" + key] = round(min(times), 3)
                built[key] = SceneData(ref.tris, idx, nodes, ref.materials, ref.tex_desc, ref.tex_data)
            ctx.setTuning(bvh_tri_cost=100)
            variants = [("reference_sbvh", ref), ("gpu_lbvh", built["fast"]), ("gpu_ploc", built["ploc"]), ("gpu_ploc_opt", built["ploc_opt"])] + [("gpu_" + k, v) for k, v in built.items() if "_tc" in k]
            variants = variants + [("reference_sbvh_again", ref)]  # run-to-run spread of the yardstick itself
            for label, sc in variants:
                depth, leaves, sah = validate_bvh(sc.nodes, sc.indices, sc.tris, unique_refs=not label.startswith("reference_sbvh"))
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
                row[label] = dict(nodes=len(sc.nodes), references=len(sc.indices), depth=depth, leaves=leaves, sah_cost=round(sah, 2),
                                  mrays_per_s=round((st.extensionRays + st.shadowRays) / ms / 1e3, 1))
        row["throughput_ratio_lbvh"] = round(row["gpu_lbvh"]["mrays_per_s"] / row["reference_sbvh"]["mrays_per_s"], 3)
        row["throughput_ratio_ploc"] = round(row["gpu_ploc"]["mrays_per_s"] / row["reference_sbvh"]["mrays_per_s"], 3)
        row["throughput_ratio_ploc_opt"] = round(row["gpu_ploc_opt"]["mrays_per_s"] / row["reference_sbvh"]["mrays_per_s"], 3)
        for k in list(row):
            if k.startswith("gpu_ploc_tc") or k.startswith("gpu_ploc_opt_tc"):
                row["throughput_ratio_" + k[4:]] = round(row[k]["mrays_per_s"] / row["reference_sbvh"]["mrays_per_s"], 3)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
