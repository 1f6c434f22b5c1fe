"""ctypes binding of libfluctus_b200.so (C ABI: include/fluctus_b200.h).  Fails loudly when the library is absent:
there is no Python or CPU implementation of the path behind it."""
import ctypes as C
import os

from .structs import QueueCounters, RenderParams, RenderStats64

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FLX_LIB_PATH") or os.path.join(_HERE, "libfluctus_b200.so")  # FLX_LIB_PATH: an experiment build (csrc/build.py)

_P = C.c_void_p
_SIGNATURES = {
    # name: (restype, argtypes)
    "flx_version": (C.c_char_p, []),
    "flx_last_error": (C.c_char_p, [_P]),
    "flx_create": (C.c_int, [C.c_int, C.c_uint32, C.POINTER(_P)]),
    "flx_destroy": (None, [_P]),
    "flx_upload_scene": (C.c_int, [_P, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_uint32, _P, C.c_size_t]),
    "flx_build_bvh": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, C.c_int, _P, C.c_uint32, C.POINTER(C.c_uint32), _P, C.POINTER(C.c_float)]),
    "flx_upload_envmap": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "flx_resize": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "flx_update_params": (C.c_int, [_P, C.POINTER(RenderParams)]),
    "flx_enqueue_reset": (C.c_int, [_P]),
    "flx_enqueue_raygen": (C.c_int, [_P]),
    "flx_enqueue_extrays": (C.c_int, [_P]),
    "flx_enqueue_shadowrays": (C.c_int, [_P]),
    "flx_enqueue_logic": (C.c_int, [_P, C.c_int]),
    "flx_enqueue_materials": (C.c_int, [_P]),
    "flx_enqueue_postprocess": (C.c_int, [_P]),
    "flx_enqueue_mk_reset": (C.c_int, [_P]),
    "flx_enqueue_mk_raygen": (C.c_int, [_P]),
    "flx_enqueue_mk_next_vertex": (C.c_int, [_P]),
    "flx_enqueue_mk_sample_bsdf": (C.c_int, [_P]),
    "flx_enqueue_mk_splat": (C.c_int, [_P]),
    "flx_enqueue_mk_splat_preview": (C.c_int, [_P]),
    "flx_render_single": (C.c_int, [_P, C.c_uint32]),
    "flx_read_preview": (C.c_int, [_P, _P, C.c_size_t]),
    "flx_set_denoiser": (C.c_int, [_P, C.c_int]),
    "flx_read_denoiser_aov": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_size_t]),
    "flx_enqueue_clear_queues": (C.c_int, [_P]),
    "flx_enqueue_get_counters": (C.c_int, [_P, C.POINTER(QueueCounters)]),
    "flx_finish": (C.c_int, [_P]),
    "flx_update_pixel_index": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "flx_reset_pixel_index": (C.c_int, [_P]),
    "flx_num_tasks": (C.c_uint32, [_P]),
    "flx_render": (C.c_int, [_P, C.c_uint32]),
    "flx_render_timed": (C.c_int, [_P, C.c_uint32, C.POINTER(C.c_float)]),
    "flx_timer_begin": (C.c_int, [_P]),
    "flx_timer_end": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "flx_set_tuning": (C.c_int, [_P, C.c_int, C.c_int]),
    "flx_set_counting": (C.c_int, [_P, C.c_int]),
    "flx_get_trace_counts": (C.c_int, [_P, _P, _P]),
    "flx_reset_stats": (C.c_int, [_P]),
    "flx_get_stats": (C.c_int, [_P, C.POINTER(RenderStats64)]),
    "flx_set_profiling": (C.c_int, [_P, C.c_int]),
    "flx_get_kernel_ms": (C.c_int, [_P, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]),
    "flx_read_pixels": (C.c_int, [_P, _P, C.c_size_t]),
    "flx_read_tasks": (C.c_int, [_P, _P]),
    "flx_read_traversal_layout": (C.c_int, [_P, _P, C.POINTER(C.c_uint32), _P, C.POINTER(C.c_uint32), C.POINTER(C.c_int32)]),
    "flx_write_tasks": (C.c_int, [_P, _P]),
    "flx_read_queue": (C.c_int, [_P, C.c_int, _P, C.c_uint32]),
    "flx_write_queue": (C.c_int, [_P, C.c_int, _P, C.c_uint32]),
    "flx_write_counters": (C.c_int, [_P, C.POINTER(QueueCounters)]),
    "flx_set_tile": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32]),
    "flx_tile_pixels": (C.c_uint32, [_P]),
    "flx_comm_unique_id": (C.c_int, [_P]),
    "flx_comm_init": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "flx_gather_pixels": (C.c_int, [_P, C.c_int, _P]),
    "flx_read_gathered": (C.c_int, [_P, C.c_int, _P, C.c_size_t]),
    "flx_comm_destroy": (C.c_int, [_P]),
    "flx_device_bytes": (C.c_size_t, [_P]),
    "flx_host_alloc": (C.c_int, [C.POINTER(_P), C.c_size_t]),
    "flx_host_free": (None, [_P]),
    "flx_io_last_error": (C.c_char_p, []),
    "flx_save_image": (C.c_int, [_P, C.c_char_p]),
    "flx_checkpoint_save": (C.c_int, [_P, C.c_char_p]),
    "flx_checkpoint_load": (C.c_int, [_P, C.c_char_p]),
    "flx_write_image": (C.c_int, [C.c_char_p, _P, C.c_uint32, C.c_uint32]),
    "flx_scene_load": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "flx_scene_free": (None, [_P]),
    "flx_scene_num_triangles": (C.c_uint32, [_P]),
    "flx_scene_num_materials": (C.c_uint32, [_P]),
    "flx_scene_num_textures": (C.c_uint32, [_P]),
    "flx_scene_triangles": (_P, [_P]),
    "flx_scene_materials": (_P, [_P]),
    "flx_scene_texture_name": (C.c_char_p, [_P, C.c_uint32]),
    "flx_image_load": (C.c_int, [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(_P)]),
    "flx_image_free": (None, [_P]),
    "flx_pack_textures": (C.c_int, [_P, _P, _P, C.c_uint32, _P, _P, C.POINTER(C.c_size_t)]),
    "flx_hierarchy_export": (C.c_int, [C.c_char_p, _P, C.c_uint32, _P, C.c_uint32]),
    "flx_hierarchy_import": (C.c_int, [C.c_char_p, _P, C.POINTER(C.c_uint32), _P, C.POINTER(C.c_uint32)]),
    "flx_envmap_load": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "flx_envmap_from_rgb": (C.c_int, [_P, C.c_int32, C.c_int32, C.POINTER(_P)]),
    "flx_envmap_free": (None, [_P]),
    "flx_envmap_width": (C.c_int32, [_P]),
    "flx_envmap_height": (C.c_int32, [_P]),
    "flx_envmap_rgb": (_P, [_P]),
    "flx_envmap_prob": (_P, [_P]),
    "flx_envmap_alias": (_P, [_P]),
    "flx_envmap_pdf": (_P, [_P]),
}
EXPORTS = tuple(_SIGNATURES)

_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a). fluctus_b200 has no fallback implementation." % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = the library does not match include/fluctus_b200.h
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib
