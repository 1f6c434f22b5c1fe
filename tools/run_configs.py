#!/usr/bin/env python
"""Throughput of the BASELINE.json configurations at full size on one GPU (run on the GPU box). One JSON line each."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CONFIGS = {  # name: (scene, W, H, NUM_TASKS)  -- SURVEY 8(d)
    "C2": ("conference", 1280, 720, 1 << 20),
    "C3": ("country_kitchen", 1920, 1080, 1 << 21),
    "C4": ("luxball", 1920, 1080, 1 << 21),
    "metric": ("conference", 1920, 1080, 1 << 21),
    "C5-tile": ("conference", 3840, 2160, 1 << 20),
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="C2,C3,C4,metric")
    ap.add_argument("--iters", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--tune", default="", help="comma-separated key=value tuning knobs (CLContext.setTuning)")
    a = ap.parse_args()
    from bench import a_ext_bytes, a_shadow_bytes, measured_peaks
    from bench_configs import ENV_MAPS, params_for
    from fluctus_b200 import CLContext, EnvMapData, SceneData, Tracer
    peak, _ = measured_peaks()
    for name in a.configs.split(","):
        scene_name, W, H, N = CONFIGS[name]
        scene = SceneData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", scene_name + ".bin"))
        params = params_for(scene_name, scene, W, H)
        with CLContext(N) as ctx:
            if a.tune:
                ctx.setTuning(**{k: int(v) for k, v in (kv.split("=") for kv in a.tune.split(","))})
            ctx.uploadSceneData(scene)
            if scene_name in ENV_MAPS:
                ctx.createEnvMap(EnvMapData.load_blob(os.path.join(ROOT, "oracle", "_ref", "scenes", ENV_MAPS[scene_name] + ".env.bin")))
            ctx.setupPixelStorage(W, H)
            tr = Tracer(ctx, params)
            tr.start()
            ctx.render(a.warmup)
            ctx.finishQueue()
            ctx.resetStats()
            ctx.setProfiling(True)
            ms = ctx.renderTimed(a.iters)
            ctx.setProfiling(False)
            st = ctx.getStats()
            perf = ctx.checkTracingPerf()
            ctx.setCounting(True)
            ctx.render(8)
            counts = ctx.getTraceCounts()
            ctx.setCounting(False)
            ae, pe = a_ext_bytes(counts["ext"])
            ash, ps = a_shadow_bytes(counts["shadow"])
            ext_ms, sh_ms = perf["extrays"][0], perf["shadowrays"][0]
            pix = ctx.readPixels()
            row = dict(config=name, scene=scene_name, width=W, height=H, num_tasks=N, max_bounces=params.maxBounces, iterations=a.iters,
                       mrays_per_s=round((st.extensionRays + st.shadowRays) / ms / 1e3, 1), ms_per_iteration=round(ms / a.iters, 4),
                       ext_mrays_per_s=round(st.extensionRays / ext_ms / 1e3, 1), shadow_mrays_per_s=round(st.shadowRays / max(sh_ms, 1e-9) / 1e3, 1),
                       A_ext_bytes=round(ae, 1), per_ray_ext=pe, A_shadow_bytes=round(ash, 1), per_ray_shadow=ps,
                       ext_roofline_frac=round(ae * st.extensionRays / (ext_ms * 1e-3) / 1e9 / peak, 4),
                       shadow_roofline_frac=round(ash * st.shadowRays / (max(sh_ms, 1e-9) * 1e-3) / 1e9 / peak, 4),
                       shadow_per_ext=round(st.shadowRays / max(st.extensionRays, 1), 3),
                       kernel_ms_per_iteration={k: round(v[0] / a.iters, 4) for k, v in perf.items() if v[1]},
                       samples=int(pix[:, 3].sum()), nonfinite_pixels=int((~__import__("numpy").isfinite(pix)).any(axis=1).sum()))
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
