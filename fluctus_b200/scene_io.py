"""File input through the C ABI (flx_scene_load / flx_envmap_load, include/fluctus_b200.h): OBJ + MTL, ASCII PLY and
Radiance RGBE in, the arrays uploadSceneData / createEnvMap take out -- the reference's Scene::loadModel and
EnvironmentMap (src/scene.cpp:52-92, src/envmap.cpp:9-114) without the reference.  Nothing is parsed in Python."""
import ctypes as C

import numpy as np

from . import _lib
from .scene import EnvMapData
from .structs import MATERIAL_DTYPE, TRIANGLE_DTYPE


class LoadedModel:
    """Triangles (file order), materials (0 = the reference's default material) and texture names (first-use order,
    relative to the model's folder); the hierarchy is built separately (CLContext.buildBVH, or any builder that emits
    the reference's Node[] format)."""

    def __init__(self, tris, materials, texture_names):
        self.tris, self.materials, self.texture_names = tris, materials, texture_names


def _io_error(lib, what):
    msg = lib.flx_io_last_error()
    from .clcontext import FluctusError
    return FluctusError("%s failed: %s" % (what, msg.decode() if msg else "?"))


def load_model(path):
    lib = _lib.load()
    h = C.c_void_p()
    if lib.flx_scene_load(str(path).encode(), C.byref(h)) != 0:
        raise _io_error(lib, "flx_scene_load(%s)" % path)
    try:
        nt, nm, nx = lib.flx_scene_num_triangles(h), lib.flx_scene_num_materials(h), lib.flx_scene_num_textures(h)
        tris = np.frombuffer(C.string_at(lib.flx_scene_triangles(h), nt * 160), TRIANGLE_DTYPE).copy()
        mats = np.frombuffer(C.string_at(lib.flx_scene_materials(h), nm * 80), MATERIAL_DTYPE).copy()
        names = [lib.flx_scene_texture_name(h, i).decode() for i in range(nx)]
    finally:
        lib.flx_scene_free(h)
    return LoadedModel(tris, mats, names)


def _env_out(lib, h):
    try:
        w, hh = lib.flx_envmap_width(h), lib.flx_envmap_height(h)
        n = w * hh
        rgb = np.frombuffer(C.string_at(lib.flx_envmap_rgb(h), n * 12), np.float32).reshape(hh, w, 3).copy()
        prob = np.frombuffer(C.string_at(lib.flx_envmap_prob(h), n * 4), np.float32).copy()
        alias = np.frombuffer(C.string_at(lib.flx_envmap_alias(h), n * 4), np.int32).copy()
        pdf = np.frombuffer(C.string_at(lib.flx_envmap_pdf(h), n * 4), np.float32).copy()
    finally:
        lib.flx_envmap_free(h)
    return EnvMapData(rgb, prob, alias, pdf)


def load_envmap(path):
    lib = _lib.load()
    h = C.c_void_p()
    if lib.flx_envmap_load(str(path).encode(), C.byref(h)) != 0:
        raise _io_error(lib, "flx_envmap_load(%s)" % path)
    return _env_out(lib, h)


def envmap_from_rgb(rgb):
    lib = _lib.load()
    rgb = np.ascontiguousarray(rgb, np.float32)
    hh, w = rgb.shape[:2]
    h = C.c_void_p()
    if lib.flx_envmap_from_rgb(rgb.ctypes.data_as(C.c_void_p), w, hh, C.byref(h)) != 0:
        raise _io_error(lib, "flx_envmap_from_rgb")
    return _env_out(lib, h)


def write_image(path, rgba, width, height):
    """CLContext::saveImage's conversions on a host RGBA float buffer (row 0 = bottom row): '*.hdr' -> RGBE of rgb / a,
    anything else -> 8-bit PNG of clamp01(rgb) (flx_write_image)."""
    lib = _lib.load()
    rgba = np.ascontiguousarray(rgba, np.float32)
    assert rgba.size == width * height * 4
    if lib.flx_write_image(str(path).encode(), rgba.ctypes.data_as(C.c_void_p), int(width), int(height)) != 0:
        raise _io_error(lib, "flx_write_image(%s)" % path)


def export_hierarchy(path, nodes, indices):
    """BVH::exportTo (src/bvh.cpp:174-192): the reference's hierarchy cache file, with the true node count."""
    lib = _lib.load()
    nodes, indices = np.ascontiguousarray(nodes), np.ascontiguousarray(indices, np.uint32)
    if lib.flx_hierarchy_export(str(path).encode(), nodes.ctypes.data_as(C.c_void_p), len(nodes), indices.ctypes.data_as(C.c_void_p), len(indices)) != 0:
        raise _io_error(lib, "flx_hierarchy_export(%s)" % path)


def import_hierarchy(path):
    """BVH::importFrom (src/bvh.cpp:102-152) -> (nodes, indices); reads files written by the reference or by export_hierarchy."""
    from .structs import NODE_DTYPE
    lib = _lib.load()
    nn, ni = C.c_uint32(), C.c_uint32()
    if lib.flx_hierarchy_import(str(path).encode(), None, C.byref(nn), None, C.byref(ni)) != 0:
        raise _io_error(lib, "flx_hierarchy_import(%s)" % path)
    nodes, indices = np.zeros(nn.value, NODE_DTYPE), np.zeros(ni.value, np.uint32)
    if lib.flx_hierarchy_import(str(path).encode(), nodes.ctypes.data_as(C.c_void_p), C.byref(nn), indices.ctypes.data_as(C.c_void_p), C.byref(ni)) != 0:
        raise _io_error(lib, "flx_hierarchy_import(%s)" % path)
    return nodes, indices


def load_model_with_textures(path):
    """Everything flx_upload_scene needs except the hierarchy, from files only: the model (flx_scene_load), its textures decoded by the
    library (flx_image_load: PNG and JPEG) and packed like CLContext::packTextures (flx_pack_textures).  Texture names are relative
    to the model's folder (src/scene.cpp:281-295).  Returns (LoadedModel, tex_desc, tex_data)."""
    import os
    model = load_model(path)
    folder = os.path.dirname(os.path.abspath(str(path)))
    images = [load_image(os.path.join(folder, name.replace("\\", "/"))) for name in model.texture_names]
    desc, blob = pack_textures(images)
    return model, desc, blob


def load_image(path):
    """PNG or JPEG -> (h, w, 4) uint8, row 0 = bottom row, the reference's in-memory texture form (flx_image_load)."""
    lib = _lib.load()
    w, h, ptr = C.c_uint32(), C.c_uint32(), C.c_void_p()
    if lib.flx_image_load(str(path).encode(), C.byref(w), C.byref(h), C.byref(ptr)) != 0:
        raise _io_error(lib, "flx_image_load(%s)" % path)
    try:
        return np.frombuffer(C.string_at(ptr, w.value * h.value * 4), np.uint8).reshape(h.value, w.value, 4).copy()
    finally:
        lib.flx_image_free(ptr)


def pack_textures(images):
    """CLContext::packTextures through the C ABI: list of (h, w, 4) uint8 arrays -> (descriptors, blob)."""
    from .structs import TEXDESC_DTYPE
    lib = _lib.load()
    images = [np.ascontiguousarray(im, np.uint8) for im in images]
    n = len(images)
    ptrs = (C.c_void_p * max(n, 1))(*[im.ctypes.data for im in images])
    ws = (C.c_uint32 * max(n, 1))(*[im.shape[1] for im in images])
    hs = (C.c_uint32 * max(n, 1))(*[im.shape[0] for im in images])
    size = C.c_size_t()
    if lib.flx_pack_textures(ptrs, ws, hs, n, None, None, C.byref(size)) != 0:
        raise _io_error(lib, "flx_pack_textures")
    desc, blob = np.zeros(n, TEXDESC_DTYPE), np.zeros(size.value, np.uint8)
    if lib.flx_pack_textures(ptrs, ws, hs, n, desc.ctypes.data_as(C.c_void_p), blob.ctypes.data_as(C.c_void_p), C.byref(size)) != 0:
        raise _io_error(lib, "flx_pack_textures")
    return desc, blob
