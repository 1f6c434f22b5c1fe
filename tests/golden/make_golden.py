#!/usr/bin/env python
"""make_golden.py -- generates the committed golden vectors by running the REFERENCE'S OWN wavefront kernels, compiled
for the host (oracle/_ref/libfluctus_ref.so, built by oracle/build_ref.py from /root/reference/src/wf_*.cl), serially.

Each fixture holds the complete path state (64 x N uint32 slots) after the prologue and after K iterations, the
radiance accumulator, the post-processed preview (mk_postprocess.cl) and the ray counts.  Inputs are regenerated deterministically by build_case(); only the
reference-derived teapot scene (reference asset assets/teapot.ply through the reference's PLY import and SBVH builder)
is stored inside its fixture, because neither exists on the GPU box or in a fresh clone.

    python tests/golden/make_golden.py          # needs /root/reference (or FLX_REFERENCE_DIR)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

CASES = ("room_mixed_separate", "room_env_mis", "teapot_c1")
# the microkernel integrator (src/mk_*.cl) through Tracer::renderSingle's loop: (case, scene/params of which wavefront case, spp)
MK_CASES = {"mk_room_env_mis": ("room_env_mis", 3), "mk_room_mixed": ("room_mixed_separate", 3)}


def run_mk_case(ctx, name, blob_loader=None):
    """Runs Tracer::renderSingle (src/tracer.cpp:112-150) on `ctx` and returns what the fixture stores: the path state after
    the first sample's first nextVertex + sampleBsdf, the state after the last splat, accumulator, preview, statistics."""
    from fluctus_b200 import Tracer
    from parity_util import mk_stats, setup_context
    base, spp = MK_CASES[name]
    scene, params, env, n, _ = build_case(base, blob_loader)
    tr = setup_context(ctx, scene, params, env)
    ctx.resetStats()
    ctx.updateParams(params)
    ctx.enqueueResetKernel(params)
    out = {}
    for s in range(spp):
        ctx.enqueueRayGenKernel(params)
        for bounce in range(params.maxBounces + 1):
            ctx.enqueueNextVertexKernel(params)
            ctx.enqueueBsdfSampleKernel(params)
            if s == 0 and bounce == 0:
                ctx.finishQueue()
                out["tasks_first_bounce"] = ctx.readTasks()
        ctx.enqueueSplatKernel(params)
        ctx.enqueuePostprocessKernel(params)
        ctx.finishQueue()
    out.update(tasks_end=ctx.readTasks(), pixels=ctx.readPixels(), preview=ctx.readPreview(), stats=np.array(mk_stats(ctx), np.int64),
               n_live=np.array([min(params.width * params.height, n)]))
    return out


def build_case(name, blob_loader=None):
    """-> (scene, params, env or None, num_tasks, iterations)"""
    from fluctus_b200 import EnvMapData, SceneData, make_params
    from fluctus_b200.scene import build_bvh, make_room_scene, room_params
    if name == "room_mixed_separate":
        scene = make_room_scene(materials="mixed", textured=True, n_blobs=8)
        return scene, room_params(scene, 48, 32, max_bounces=6, separate_queues=True), None, 2048, 12
    if name == "room_env_mis":
        scene = make_room_scene(materials="mixed", textured=True)
        keep = np.ones(len(scene.tris), bool)
        keep[2:4] = False
        keep[6:8] = False
        tris = scene.tris[keep]
        nodes, indices = build_bvh(tris)
        scene = SceneData(tris, indices, nodes, scene.materials, scene.tex_desc, scene.tex_data)
        rng = np.random.default_rng(3)
        rgb = rng.uniform(0.0, 0.4, size=(16, 32, 3)).astype(np.float32)
        rgb[3:5, 5:8] += 25.0
        env = EnvMapData.from_rgb(rgb)
        return scene, room_params(scene, 40, 24, max_bounces=5, separate_queues=False, use_env_map=True, use_area_light=True, env_map_strength=2.0), env, 960, 12
    if name == "teapot_c1":  # BASELINE config C1 at 64x64: teapot.ply, 2 bounces, default camera and light (tracer.cpp:760-797)
        fix = os.path.join(HERE, "teapot_c1.npz")
        if os.path.exists(fix):
            z = np.load(fix)
            from fluctus_b200.structs import MATERIAL_DTYPE, NODE_DTYPE, TRIANGLE_DTYPE
            scene = SceneData(z["scene_tris"].view(TRIANGLE_DTYPE).reshape(-1), z["scene_indices"], z["scene_nodes"].view(NODE_DTYPE).reshape(-1),
                              z["scene_materials"].view(MATERIAL_DTYPE).reshape(-1), name="teapot")
        else:
            scene = SceneData.load_blob(blob_loader("teapot") if blob_loader else os.path.join(ROOT, "oracle", "_ref", "scenes", "teapot.bin"))
        cam = dict(pos=(0, 1, 3.5), dir=(0, 0, -1), right=(1, 0, 0), up=(0, 1, 0), fov=60.0)
        return scene, make_params(64, 64, cam, scene.world_radius, len(scene.tris), max_bounces=2), None, 4096, 16
    raise KeyError(name)


def build_ref_dir():
    from oracle import build_ref
    return build_ref.reference_dir()


IO_FIXTURES = (("obj", "tricky.obj"), ("ply", "tricky.ply"), ("ply", "plain.ply"), ("env", "small_rle.hdr"))


def make_io_expected():
    """Runs the REFERENCE'S loader code (oracle/_ref/scene_tool: its vendored tinyobjloader + the scene.cpp conversion, its
    PLY reader, its RGBE reader and importance tables) on the small input files under tests/golden/io/ and stores what it
    produced; tests/test_scene_io_cpu.py holds flx_scene_load / flx_envmap_load to these bytes."""
    import struct
    import subprocess
    import tempfile
    from fluctus_b200.structs import MATERIAL_DTYPE, TRIANGLE_DTYPE
    tool = os.path.join(ROOT, "oracle", "_ref", "scene_tool")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for mode, name in IO_FIXTURES:
            blob = os.path.join(tmp, name + ".bin")
            subprocess.run([tool, mode, os.path.join(HERE, "io", name), blob], check=True, capture_output=True)
            buf = open(blob, "rb").read()
            key = name.replace(".", "_")
            if mode == "env":
                magic, w, h = struct.unpack_from("<3I", buf, 0)
                n, off = w * h, 12
                out[key + "_size"] = np.array([w, h])
                out[key + "_rgb"] = np.frombuffer(buf, np.float32, n * 3, off).copy(); off += n * 12
                out[key + "_prob"] = np.frombuffer(buf, np.float32, n, off).copy(); off += n * 4
                out[key + "_alias"] = np.frombuffer(buf, np.int32, n, off).copy(); off += n * 4
                out[key + "_pdf"] = np.frombuffer(buf, np.float32, n, off).copy()
                continue
            magic, nt, ni, nn, nm, ntex = struct.unpack_from("<6I", buf, 0)
            off = 24
            out[key + "_tris"] = np.frombuffer(buf, np.uint8, nt * 160, off).copy(); off += nt * 160 + ni * 4 + nn * 48
            out[key + "_materials"] = np.frombuffer(buf, np.uint8, nm * 80, off).copy(); off += nm * 80
            names = []
            for _ in range(ntex):
                (ln,) = struct.unpack_from("<I", buf, off); off += 4
                names.append(buf[off:off + ln].decode()); off += ln
            out[key + "_textures"] = np.array(names, dtype="U128")
    # the reference's hierarchy cache file for the teapot, written by the reference's own BVH::exportTo (src/bvh.cpp:174-192)
    subprocess.run([tool, "cache-export", "ply", os.path.join(build_ref_dir(), "assets", "teapot.ply"), os.path.join(HERE, "io", "teapot_hierarchy.bin")],
                   check=True, capture_output=True)
    fix = os.path.join(HERE, "io", "expected.npz")
    np.savez_compressed(fix, **out)
    print("io fixtures:", os.path.getsize(fix), "bytes")


def main():
    from oracle import build_ref, make_scenes
    from oracle.oracle_host import RefContext
    from parity_util import setup_context
    build_ref.build()
    make_scenes.build(["teapot"])
    for name in CASES:
        fix = os.path.join(HERE, name + ".npz")
        if name == "teapot_c1" and os.path.exists(fix):
            os.remove(fix)  # rebuild the scene from the reference asset, not from the old fixture
        scene, params, env, n, iters = build_case(name)
        ctx = RefContext(n)
        tr = setup_context(ctx, scene, params, env)
        tr.start()
        out = dict(tasks_start=ctx.readTasks())
        for _ in range(iters):
            tr.iterate()
        out.update(tasks_end=ctx.readTasks(), pixels=ctx.readPixels(), preview=ctx.readPreview(), stats=np.array([tr.stats[k] for k in ("primaryRays", "extensionRays", "shadowRays")], np.int64))
        if name == "teapot_c1":
            out.update(scene_tris=scene.tris.view(np.uint8), scene_indices=scene.indices, scene_nodes=scene.nodes.view(np.uint8), scene_materials=scene.materials.view(np.uint8))
        np.savez_compressed(fix, **out)
        print(name, os.path.getsize(fix), "bytes;", dict(zip(("primary", "extension", "shadow"), out["stats"])))
    make_io_expected()
    for name, (base, spp) in MK_CASES.items():
        n = build_case(base)[3]
        out = run_mk_case(RefContext(n), name)
        fix = os.path.join(HERE, name + ".npz")
        np.savez_compressed(fix, **out)
        print(name, os.path.getsize(fix), "bytes;", dict(zip(("primary", "extension", "shadow", "samples"), out["stats"])))


if __name__ == "__main__":
    main()
