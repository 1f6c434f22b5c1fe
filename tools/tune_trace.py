#!/usr/bin/env python
"""Sweep the traversal tuning knobs on a benchmark scene and print per-kernel device time (run on the GPU box)."""
import argparse
import itertools
import json
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
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--variants", default="0,1")
    ap.add_argument("--thresholds", default="8,16,20,24,28,32")
    ap.add_argument("--blocks", default="0")
    ap.add_argument("--tops", default="1023,2047,3071")
    ap.add_argument("--inner-mins", default="8")
    ap.add_argument("--logic-blocks", default="3")
    ap.add_argument("--chunks", default="32")
    ap.add_argument("--overlaps", default="1")
    ap.add_argument("--ext-blocks", default="8")
    ap.add_argument("--shadow-blocks", default="10")
    ap.add_argument("--smem-stacks", default="0")
    ap.add_argument("--max-l1", default="1")
    ap.add_argument("--builder", default="", help="fast | ploc: trace through a GPU-built hierarchy instead of the scene file's SBVH")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    from bench_configs import ENV_MAPS, params_for
    scene = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", a.scene + ".bin"))
    params = params_for(a.scene, scene, a.width, a.height)
    ctx = CLContext(a.tasks)
    if a.builder:
        nodes, idx, _ = ctx.buildBVH(scene.tris, 8, a.builder)
        scene = SceneData(scene.tris, idx, nodes, scene.materials, scene.tex_desc, scene.tex_data)
    ctx.uploadSceneData(scene)
    if a.scene in ENV_MAPS:
        ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", ENV_MAPS[a.scene] + ".env.bin")))
    ctx.setupPixelStorage(a.width, a.height)
    rows = []
    combos = []
    for v in map(int, a.variants.split(",")):
        if v == 0:
            combos.append((0, 0, 0, 0, 1))
        elif v == 1:
            combos += [(v, th, b, 0, im) for th, b, im in itertools.product(map(int, a.thresholds.split(",")), map(int, a.blocks.split(",")), map(int, a.inner_mins.split(",")))]
        else:
            combos += [(v, th, 0, top, im) for th, top, im in itertools.product(map(int, a.thresholds.split(",")), map(int, a.tops.split(",")), map(int, a.inner_mins.split(",")))]
    combos = [c + (lb, ch) for c in combos for lb in map(int, a.logic_blocks.split(",")) for ch in map(int, a.chunks.split(","))]
    combos = [c + (ov,) for c in combos for ov in map(int, a.overlaps.split(","))]
    combos = [c + (eb, sb) for c in combos for eb in map(int, a.ext_blocks.split(",")) for sb in map(int, a.shadow_blocks.split(","))]
    combos = [c + (ss, ml) for c in combos for ss in map(int, a.smem_stacks.split(",")) for ml in map(int, a.max_l1.split(","))]
    for v, th, b, top, im, lb, ch, ov, eb, sb, ss, ml in combos:
        ctx.setTuning(trace_variant=v, logic_min_blocks=lb, fetch_chunk=ch, overlap_trace=ov, ext_min_blocks=eb, shadow_min_blocks=sb, smem_stack=ss, max_l1=ml)
        if v:
            ctx.setTuning(fetch_threshold=th, trace_blocks_per_sm=b, top_nodes=top, inner_min=im)
        tr = Tracer(ctx, params)
        tr.start()
        ctx.render(12)
        ctx.finishQueue()
        ctx.resetStats()
        ctx.setProfiling(True)
        ms = ctx.renderTimed(a.iters)
        ctx.setProfiling(False)
        st = ctx.getStats()
        perf = ctx.checkTracingPerf()
        row = dict(variant=v, threshold=th, blocks_per_sm=b, top_nodes=top, inner_min=im, logic_blocks=lb, fetch_chunk=ch, overlap=ov, ext_blocks=eb, shadow_blocks=sb, smem_stack=ss, max_l1=ml, ms_per_iter=ms / a.iters, mrays=(st.extensionRays + st.shadowRays) / ms / 1e3,
                   ext_ms=perf["extrays"][0] / a.iters, shadow_ms=perf["shadowrays"][0] / a.iters, logic_ms=perf["logic"][0] / a.iters,
                   mat_ms=perf["materials"][0] / a.iters, raygen_ms=perf["raygen"][0] / a.iters,
                   ext_mrays=st.extensionRays / perf["extrays"][0] / 1e3, shadow_mrays=st.shadowRays / perf["shadowrays"][0] / 1e3)
        rows.append(row)
        print(json.dumps({k: (round(x, 4) if isinstance(x, float) else x) for k, x in row.items()}), flush=True)
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
