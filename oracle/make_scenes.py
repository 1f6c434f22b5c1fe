#!/usr/bin/env python
"""make_scenes.py -- TEST INFRASTRUCTURE (oracle).

Runs oracle/_ref/scene_tool (the reference's own OBJ/PLY import, SBVH builder and env-map table code, see
oracle/ref_shim/scene_tool.cpp) on the reference's assets and writes the scene blobs the tests and bench.py load:

    oracle/_ref/scenes/{teapot,conference,luxball,country_kitchen,egyptcat}.bin  (+ .tex.npz for textured scenes)
    oracle/_ref/scenes/night.env.bin

Blobs are git-ignored (they derive from /root/reference/assets) but are NOT gpurun-ignored, so they travel to the
GPU box, where /root/reference does not exist.  Textures are decoded here once with Pillow (the reference uses DevIL,
which is not available: "texture decode parity unpinned") so that oracle and CUDA path consume the same bytes.
"""
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
OUT = os.path.join(HERE, "_ref", "scenes")
TOOL = os.path.join(HERE, "_ref", "scene_tool")

SCENES = {
    "teapot": ("ply", "assets/teapot.ply"),
    "conference": ("obj", "assets/conference/conference.obj"),
    "luxball": ("obj", "assets/luxball/luxball.obj"),
    "country_kitchen": ("obj", "assets/country_kitchen/Country-Kitchen.obj"),
    "egyptcat": ("obj", "assets/egyptcat/egyptcat.obj"),  # first scene of the reference's own benchmark protocol (tracer.cpp:384-389)
}
ENVMAPS = {"night": "assets/env_maps/night.hdr"}
# Asset FILES that travel as they are (oracle/_ref/assets/, git-ignored like the blobs): the inputs of the "from files only" tests --
# model + materials + JPEG / PNG textures + environment map through the library's own loaders and decoders, no blob, no Pillow.
ASSET_DIRS = {"country_kitchen": "assets/country_kitchen", "env_maps": "assets/env_maps"}
ASSETS_OUT = os.path.join(HERE, "_ref", "assets")


def build(names=None, force=False):
    from oracle import build_ref
    from fluctus_b200.scene import SceneData, pack_textures
    ref = build_ref.reference_dir()
    if not build_ref.available():
        raise RuntimeError("reference tree not found at %s" % ref)
    build_ref.build()
    os.makedirs(OUT, exist_ok=True)
    made = []
    for name, (mode, rel) in SCENES.items():
        if names and name not in names:
            continue
        out = os.path.join(OUT, name + ".bin")
        if force or not os.path.exists(out):
            r = subprocess.run([TOOL, mode, os.path.join(ref, rel), out], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("scene_tool failed on %s:\n%s" % (rel, r.stderr[-2000:]))
        made.append(out)
        side = os.path.join(OUT, name + ".tex.npz")
        if force or not os.path.exists(side):
            try:
                scene = SceneData.load_blob(out, texture_root=os.path.dirname(os.path.join(ref, rel)))
            except FileNotFoundError:
                raise
            if getattr(scene, "texture_names", None):
                np.savez_compressed(side, desc=scene.tex_desc.view(np.uint32).reshape(-1, 3), data=scene.tex_data)
                made.append(side)
    import shutil
    for name, rel in ASSET_DIRS.items():
        if names and name not in names:
            continue
        dst = os.path.join(ASSETS_OUT, name)
        if force or not os.path.isdir(dst):
            shutil.rmtree(dst, ignore_errors=True)
            if name == "env_maps":  # only the map the benchmark configuration uses
                os.makedirs(dst)
                shutil.copy(os.path.join(ref, rel, "night.hdr"), dst)
            else:
                shutil.copytree(os.path.join(ref, rel), dst)
        made.append(dst)
    for name, rel in ENVMAPS.items():
        if names and name not in names:
            continue
        out = os.path.join(OUT, name + ".env.bin")
        if force or not os.path.exists(out):
            r = subprocess.run([TOOL, "env", os.path.join(ref, rel), out], capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("scene_tool failed on %s:\n%s" % (rel, r.stderr[-2000:]))
        made.append(out)
    return made


if __name__ == "__main__":
    for p in build(sys.argv[1:] and [a for a in sys.argv[1:] if not a.startswith("-")] or None, force="--force" in sys.argv):
        print(p, os.path.getsize(p) if os.path.isfile(p) else "(directory)")
