"""Host-side scene containers for the hot path's INPUT formats.

The path consumes what the reference uploads in CLContext::uploadSceneData / createEnvMap
(reference: src/clcontext.cpp:467-611): 160-byte triangles, a u32 index list, 48-byte BVH nodes in
DFS order (left child = self + 1, src/bvhnode.hpp:50-59), 80-byte materials, packed RGBA8 textures and the
environment map with its alias tables.  Loading OBJ files and building an SBVH are out of scope
(SURVEY 8f): real scenes arrive as blobs written by the reference's own loader/builder
(oracle/ref_shim/scene_tool.cpp).  For self-contained tests this module also has a small
procedural scene and a plain median-split BVH builder that emits the reference node format.
"""
import math
import os
import struct

import numpy as np

from .structs import (BXDF, MATERIAL_DTYPE, NODE_DTYPE, TEXDESC_DTYPE, TRIANGLE_DTYPE, RenderParams)


def _aligned(a, dtype, align=64):
    a = np.asarray(a, dtype=dtype).reshape(-1)
    raw = np.empty(a.nbytes + align, np.uint8)
    off = (-raw.ctypes.data) % align
    out = raw[off:off + a.nbytes].view(dtype)
    out[...] = a
    return out


class SceneData:
    def __init__(self, tris, indices, nodes, materials, tex_desc=None, tex_data=None, name="scene"):
        # fresh, owned, 16-byte aligned copies (CPU consumers load the 16-byte float3 members with aligned moves)
        self.tris = _aligned(tris, TRIANGLE_DTYPE)
        self.indices = _aligned(indices, np.uint32)
        self.nodes = _aligned(nodes, NODE_DTYPE)
        self.materials = _aligned(materials, MATERIAL_DTYPE)
        self.tex_desc = _aligned(tex_desc if tex_desc is not None else np.zeros(0, TEXDESC_DTYPE), TEXDESC_DTYPE)
        self.tex_data = _aligned(tex_data if tex_data is not None else np.zeros(0, np.uint8), np.uint8)
        self.name = name

    @property
    def world_radius(self):
        # Tracer::init: half the diagonal of the root box (reference: src/tracer.cpp:66-67), float32 arithmetic
        d = (self.nodes[0]["bmax"][:3].astype(np.float32) - self.nodes[0]["bmin"][:3].astype(np.float32)).astype(np.float32)
        return float(np.float32(0.5) * np.sqrt(np.float32(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]), dtype=np.float32))

    @property
    def material_types(self):
        return int(np.bitwise_or.reduce(self.materials["type"])) if len(self.materials) else 0

    def pinned(self):
        """The same scene with every array in page-locked host memory (flx_host_alloc), so uploadSceneData copies by DMA."""
        from .clcontext import pinned_copy
        out = SceneData.__new__(SceneData)
        for k in ("tris", "indices", "nodes", "materials", "tex_desc", "tex_data"):
            a = getattr(self, k)
            setattr(out, k, pinned_copy(a) if a.size else a)
        out.name = self.name
        if hasattr(self, "texture_names"):
            out.texture_names = self.texture_names
        return out

    def nbytes(self):
        return self.tris.nbytes + self.indices.nbytes + self.nodes.nbytes + self.materials.nbytes + self.tex_desc.nbytes + self.tex_data.nbytes

    # ---- blob written by oracle/ref_shim/scene_tool.cpp ("FLXS")
    @staticmethod
    def load_blob(path, texture_root=None):
        with open(path, "rb") as f:
            buf = f.read()
        magic, nt, ni, nn, nm, ntex = struct.unpack_from("<6I", buf, 0)
        if magic != 0x53584C46:
            raise ValueError("%s: not a scene blob" % path)
        off = 24
        tris = np.frombuffer(buf, TRIANGLE_DTYPE, nt, off); off += nt * 160
        indices = np.frombuffer(buf, np.uint32, ni, off); off += ni * 4
        nodes = np.frombuffer(buf, NODE_DTYPE, nn, off); off += nn * 48
        mats = np.frombuffer(buf, MATERIAL_DTYPE, nm, off); off += nm * 80
        names = []
        for _ in range(ntex):
            (ln,) = struct.unpack_from("<I", buf, off); off += 4
            names.append(buf[off:off + ln].decode()); off += ln
        scene = SceneData(tris, indices, nodes, mats, name=os.path.splitext(os.path.basename(path))[0])
        scene.texture_names = names
        if names:
            side = os.path.splitext(path)[0] + ".tex.npz"
            if os.path.exists(side):  # textures decoded once (oracle/make_scenes.py) so every consumer sees the same bytes
                z = np.load(side)
                scene.tex_desc, scene.tex_data = np.ascontiguousarray(z["desc"].view(TEXDESC_DTYPE).reshape(-1)), np.ascontiguousarray(z["data"])
            elif texture_root is not None:
                scene.tex_desc, scene.tex_data = pack_textures([os.path.join(texture_root, n) for n in names])
            else:
                raise FileNotFoundError("scene %s needs its decoded textures (%s)" % (path, side))
        return scene


def pack_textures(paths):
    """RGBA8, origin lower-left, concatenated (reference: src/texture.cpp:16-40 + ilOriginFunc(IL_ORIGIN_LOWER_LEFT) in
    src/main.cpp:69-71, packing src/clcontext.cpp:570-611).  Decoding uses Pillow instead of DevIL."""
    from PIL import Image

    desc = np.zeros(len(paths), TEXDESC_DTYPE)
    chunks, offset = [], 0
    for i, p in enumerate(paths):
        img = np.asarray(Image.open(p).convert("RGBA"), dtype=np.uint8)[::-1]  # flip to lower-left origin
        h, w = img.shape[:2]
        desc[i] = (offset, w, h)
        chunks.append(np.ascontiguousarray(img).reshape(-1))
        offset += w * h * 4
    return desc, (np.concatenate(chunks) if chunks else np.zeros(0, np.uint8))


class EnvMapData:
    def __init__(self, rgb, prob, alias, pdf):
        self.rgb = np.ascontiguousarray(rgb, np.float32)  # (h, w, 3)
        self.height, self.width = self.rgb.shape[:2]
        self.prob = np.ascontiguousarray(prob, np.float32).reshape(-1)
        self.alias = np.ascontiguousarray(alias, np.int32).reshape(-1)
        self.pdf = np.ascontiguousarray(pdf, np.float32).reshape(-1)

    @staticmethod
    def load_blob(path):  # "FLXE", written by scene_tool env
        with open(path, "rb") as f:
            buf = f.read()
        magic, w, h = struct.unpack_from("<3I", buf, 0)
        if magic != 0x45584C46:
            raise ValueError("%s: not an env-map blob" % path)
        n, off = w * h, 12
        rgb = np.frombuffer(buf, np.float32, n * 3, off).reshape(h, w, 3); off += n * 12
        prob = np.frombuffer(buf, np.float32, n, off); off += n * 4
        alias = np.frombuffer(buf, np.int32, n, off); off += n * 4
        pdf = np.frombuffer(buf, np.float32, n, off)
        return EnvMapData(rgb, prob, alias, pdf)

    @staticmethod
    def from_rgb(rgb):
        """Importance tables for an RGB lat-long image, the way EnvironmentMap::computeProbabilities does it
        (reference: src/envmap.cpp:31-114): luminance*sin(theta) -> pdf with mean 1 -> Vose alias tables built with two
        LIFO stacks (the stack order decides which table comes out)."""
        rgb = np.ascontiguousarray(rgb, np.float32)
        h, w = rgb.shape[:2]
        n = w * h
        f32 = np.float32
        # float sinTh = std::sin(PI * float(v + 0.5f) / float(height)): float32 argument, sinf
        arg = (f32(np.pi) * (np.arange(h, dtype=np.float32) + f32(0.5))).astype(np.float32) / f32(h)
        try:  # the C library's sinf, which is what std::sin(float) calls in the reference (not always correctly rounded)
            import ctypes
            libm = ctypes.CDLL("libm.so.6")
            libm.sinf.restype, libm.sinf.argtypes = ctypes.c_float, [ctypes.c_float]
            sin_th = np.array([libm.sinf(float(a)) for a in arg], np.float32)
        except OSError:
            sin_th = np.sin(arg.astype(np.float64)).astype(np.float32)
        lum = (f32(0.212671) * rgb[..., 0] + f32(0.715160) * rgb[..., 1]).astype(np.float32) + f32(0.072169) * rgb[..., 2]
        scal = (lum.astype(np.float32) * sin_th[:, None]).astype(np.float32).reshape(-1)
        I = f32(0.0)
        fn = f32(n)
        for s in scal:  # serial float accumulation, as the reference
            I = f32(I + f32(s / fn))
        pdf = (np.full(n, f32(1.0) / fn, np.float32) if I == 0 else (scal / I).astype(np.float32))
        prob = np.zeros(n, np.float32)
        alias = np.zeros(n, np.int32)
        small, large = [], []
        for i in range(n):
            (small if pdf[i] < 1.0 else large).append((f32(pdf[i]), i))
        while small and large:
            lp, li = small.pop()
            gp, gi = large.pop()
            prob[li] = lp
            alias[li] = gi
            pg = f32(f32(gp + lp) - f32(1.0))
            (small if pg < 1.0 else large).append((pg, gi))
        for _, gi in large:
            prob[gi] = 1.0
        for _, li in small:
            prob[li] = 1.0
        return EnvMapData(rgb, prob, alias, pdf)


# ------------------------------------------------------------------------------------------------ params
def _norm(v):
    v = np.asarray(v, np.float64)
    return v / np.linalg.norm(v)


def look_at(pos, target, fov=60.0, up=(0.0, 1.0, 0.0), aperture=0.0, focal_dist=0.5):
    """Camera basis from 'pos -> target' (SURVEY 8d): dir = normalize(target - pos), right = normalize(cross(dir, up)),
    up = cross(right, dir); stored as float32 like the reference's Camera (geom.h:146-155)."""
    d = _norm(np.asarray(target, np.float64) - np.asarray(pos, np.float64))
    r = _norm(np.cross(d, np.asarray(up, np.float64)))
    u = np.cross(r, d)
    return dict(pos=np.asarray(pos, np.float32), dir=d.astype(np.float32), right=r.astype(np.float32), up=u.astype(np.float32), fov=float(fov),
                apertureSize=float(aperture), focalDist=float(focal_dist))


def make_params(width, height, camera, world_radius, n_tris=0, light=None, max_bounces=8, use_env_map=False, env_map_strength=1.0,
                sample_impl=True, sample_expl=True, use_roulette=False, separate_queues=False):
    """RenderParams with the reference's defaults (src/tracer.cpp:38-52, 760-797) overridden by the arguments."""
    p = RenderParams()
    p.width, p.height, p.n_tris = int(width), int(height), int(n_tris)
    c = p.camera
    c.pos.set(camera["pos"]); c.dir.set(camera["dir"]); c.up.set(camera["up"]); c.right.set(camera["right"])
    c.fov, c.apertureSize, c.focalDist = camera["fov"], camera.get("apertureSize", 0.0), camera.get("focalDist", 0.5)
    p.ppParams.exposure, p.ppParams.tmOperator = 1.0, 2
    a = p.areaLight
    if light is None:  # Tracer::initAreaLight, src/tracer.cpp:788-797
        light = dict(pos=(1.0, 1.0, 0.0, 1.0), N=(-1.0, 0.0, 0.0, 0.0), right=(0.0, 0.0, -1.0), up=(0.0, 1.0, 0.0), size=(0.5, 0.5), E=(200.0, 200.0, 200.0))
        p.useAreaLight = 1
    else:
        p.useAreaLight = 0 if light is False else 1
        if light is False:
            light = dict(pos=(0, 0, 0), N=(0, -1, 0), right=(1, 0, 0), up=(0, 0, 1), size=(0.5, 0.5), E=(0, 0, 0))
    a.pos.set(light["pos"]); a.N.set(light["N"]); a.right.set(light["right"]); a.up.set(light["up"]); a.E.set(light["E"])
    a.size.x, a.size.y = float(light["size"][0]), float(light["size"][1])
    p.useEnvMap = 1 if use_env_map else 0
    p.envMapStrength = float(env_map_strength)
    p.maxBounces = int(max_bounces)
    p.sampleImpl, p.sampleExpl, p.useRoulette = int(bool(sample_impl)), int(bool(sample_expl)), int(bool(use_roulette))
    p.wfSeparateQueues = int(bool(separate_queues))
    p.worldRadius = float(world_radius)
    return p


# ------------------------------------------------------------------------------------------------ procedural scenes
def _material(kd=(0.64, 0.64, 0.64), ks=(0, 0, 0), ns=700.0, ni=1.8, type_=BXDF.DIFFUSE, map_kd=-1, map_ks=-1, map_n=-1):
    m = np.zeros((), MATERIAL_DTYPE)
    m["Kd"][:3], m["Ks"][:3] = kd, ks
    m["Ns"], m["Ni"], m["map_Kd"], m["map_Ks"], m["map_N"], m["type"] = ns, ni, map_kd, map_ks, map_n, type_
    return m


def _tri(p0, p1, p2, mat, n=None, uv=None):
    t = np.zeros((), TRIANGLE_DTYPE)
    p = [np.asarray(q, np.float32) for q in (p0, p1, p2)]
    if n is None:
        fn = np.cross(p[1] - p[0], p[2] - p[0]).astype(np.float64)
        ln = np.linalg.norm(fn)
        fn = (fn / ln if ln > 0 else fn).astype(np.float32)
        n = (fn, fn, fn)
    if uv is None:
        uv = ((0, 0), (1, 0), (0, 1))
    for k, key in enumerate(("v0", "v1", "v2")):
        t[key]["p"][:3] = p[k]
        t[key]["n"][:3] = n[k]
        t[key]["t"][:2] = uv[k]
    t["matId"] = mat
    return t


def build_bvh(tris, max_leaf=4):
    """Median-split BVH over triangle centroids in the reference's node format (DFS order, left = self + 1,
    `link` = rightChild for inner nodes / iStart for leaves, nPrims = 0 marks an inner node)."""
    P = np.stack([tris["v0"]["p"][:, :3], tris["v1"]["p"][:, :3], tris["v2"]["p"][:, :3]], axis=1).astype(np.float32)
    lo, hi = P.min(axis=1), P.max(axis=1)
    cen = (lo + hi) * np.float32(0.5)
    nodes, order = [], []

    def rec(ids, parent):
        me = len(nodes)
        n = np.zeros((), NODE_DTYPE)
        n["bmin"][:3], n["bmax"][:3], n["parent"] = lo[ids].min(axis=0), hi[ids].max(axis=0), parent
        nodes.append(n)
        if len(ids) <= max_leaf:
            n["link"], n["nPrims"] = len(order), len(ids)
            order.extend(int(i) for i in ids)
            return
        ext = cen[ids].max(axis=0) - cen[ids].min(axis=0)
        axis = int(np.argmax(ext))
        srt = ids[np.argsort(cen[ids, axis], kind="stable")]
        half = len(srt) // 2
        rec(srt[:half], me)
        nodes[me]["link"] = len(nodes)
        rec(srt[half:], me)

    import sys
    sys.setrecursionlimit(10000)
    rec(np.arange(len(tris)), -1)
    return np.array(nodes, NODE_DTYPE), np.array(order, np.uint32)


def make_room_scene(seed=0, n_blobs=6, detail=6, materials="diffuse", textured=False):
    """A closed room (so every path keeps bouncing) with a ceiling light gap and a few tessellated spheres.
    materials: "diffuse" (all Lambert) or "mixed" (one sphere per BSDF type of bxdf_types.h:4-11)."""
    rng = np.random.default_rng(seed)
    mats = [_material()]  # id 0 = the reference's default material (scene.cpp:13-26)
    mats.append(_material(kd=(0.7, 0.2, 0.2)))
    mats.append(_material(kd=(0.2, 0.7, 0.2)))
    mats.append(_material(kd=(0.75, 0.75, 0.75), map_kd=0 if textured else -1))
    if materials == "mixed":
        mats.append(_material(kd=(0.5, 0.4, 0.3), ks=(0.04, 0.04, 0.04), ns=200.0, ni=1.5, type_=BXDF.GLOSSY))
        mats.append(_material(ks=(0.9, 0.8, 0.6), ns=80.0, ni=0.0, type_=BXDF.GGX_ROUGH_REFLECTION))
        mats.append(_material(ks=(0.95, 0.95, 0.95), type_=BXDF.IDEAL_REFLECTION))
        mats.append(_material(ks=(0.9, 0.95, 1.0), ns=300.0, ni=1.5, type_=BXDF.GGX_ROUGH_DIELECTRIC))
        mats.append(_material(ks=(1.0, 1.0, 1.0), ni=1.5, type_=BXDF.IDEAL_DIELECTRIC))
        mats.append(_material(kd=(0.3, 0.3, 0.6), ks=(0.0, 0.0, 0.0), ns=50.0, ni=0.0, type_=BXDF.GLOSSY, map_n=1 if textured else -1))
    tris = []

    def quad(a, b, c, d, m):
        tris.append(_tri(a, b, c, m, uv=((0, 0), (1, 0), (1, 1))))
        tris.append(_tri(a, c, d, m, uv=((0, 0), (1, 1), (0, 1))))

    X, Y, Z = 1.0, 1.0, 1.0
    quad((-X, 0, -Z), (X, 0, -Z), (X, 0, Z), (-X, 0, Z), 3)          # floor (normal +y after winding? see below)
    quad((-X, 2 * Y, Z), (X, 2 * Y, Z), (X, 2 * Y, -Z), (-X, 2 * Y, -Z), 0)  # ceiling
    quad((-X, 0, -Z), (-X, 2 * Y, -Z), (X, 2 * Y, -Z), (X, 0, -Z), 0)  # back
    quad((X, 0, Z), (X, 2 * Y, Z), (-X, 2 * Y, Z), (-X, 0, Z), 0)      # front
    quad((-X, 0, Z), (-X, 2 * Y, Z), (-X, 2 * Y, -Z), (-X, 0, -Z), 1)  # left
    quad((X, 0, -Z), (X, 2 * Y, -Z), (X, 2 * Y, Z), (X, 0, Z), 2)      # right
    n_mats = len(mats)
    for b in range(n_blobs):
        c = np.array([rng.uniform(-0.6, 0.6), rng.uniform(0.3, 1.2), rng.uniform(-0.6, 0.6)])
        r = rng.uniform(0.15, 0.3)
        m = (4 + b) % n_mats if materials == "mixed" else 1 + b % 3
        if materials == "mixed" and m < 4:
            m = 4 + (b % (n_mats - 4))
        nu, nv = detail * 2, detail
        def pt(i, j):
            th, ph = math.pi * j / nv, 2 * math.pi * i / nu
            d = np.array([math.sin(th) * math.cos(ph), math.cos(th), math.sin(th) * math.sin(ph)])
            return c + r * d, d
        for j in range(nv):
            for i in range(nu):
                (p00, n00), (p10, n10), (p01, n01), (p11, n11) = pt(i, j), pt(i + 1, j), pt(i, j + 1), pt(i + 1, j + 1)
                uv = lambda i_, j_: (i_ / nu, j_ / nv)
                if j != 0:
                    tris.append(_tri(p00, p10, p11, m, n=(n00, n10, n11), uv=(uv(i, j), uv(i + 1, j), uv(i + 1, j + 1))))
                if j != nv - 1:
                    tris.append(_tri(p00, p11, p01, m, n=(n00, n11, n01), uv=(uv(i, j), uv(i + 1, j + 1), uv(i, j + 1))))
    tris = np.array(tris, TRIANGLE_DTYPE)
    nodes, indices = build_bvh(tris)
    tex_desc = tex_data = None
    if textured:
        t0 = rng.integers(0, 256, size=(16, 16, 4), dtype=np.uint8)
        t1 = rng.integers(96, 160, size=(8, 8, 4), dtype=np.uint8)
        t1[..., 2] = 255
        tex_desc = np.zeros(2, TEXDESC_DTYPE)
        tex_desc[0] = (0, 16, 16)
        tex_desc[1] = (16 * 16 * 4, 8, 8)
        tex_data = np.concatenate([t0.reshape(-1), t1.reshape(-1)])
    return SceneData(tris, indices, nodes, np.array(mats, MATERIAL_DTYPE), tex_desc, tex_data, name="room_%s" % materials)


def room_params(scene, width, height, max_bounces=4, separate_queues=False, use_env_map=False, use_area_light=True, **kw):
    cam = look_at((0.0, 1.0, 0.95), (0.0, 0.9, -0.2), fov=70.0)
    light = dict(pos=(0.0, 1.98, 0.0), N=(0.0, -1.0, 0.0), right=(1.0, 0.0, 0.0), up=(0.0, 0.0, 1.0), size=(0.3, 0.3), E=(60.0, 60.0, 60.0))
    return make_params(width, height, cam, scene.world_radius, n_tris=len(scene.tris), light=light if use_area_light else False, max_bounces=max_bounces,
                       separate_queues=separate_queues, use_env_map=use_env_map, **kw)
