"""CLContext -- host-side mirror of the reference's device context for the wavefront path.

Same method names, argument meaning and error behaviour as the reference class (src/clcontext.hpp:26-211):
methods enqueue work asynchronously on one in-order queue, `finishQueue` is the only synchronisation point, and
failures raise (the reference throws std::runtime_error through clt::check, ext/CLT/src/utils.cpp:22-29).
Everything is forwarded to the C ABI in include/fluctus_b200.h; nothing is computed here.
"""
import ctypes as C

import numpy as np

from . import _lib
from .structs import QUEUE_NAMES, QueueCounters, RenderParams, RenderStats64


class FluctusError(RuntimeError):
    pass


KERNEL_IDS = {"reset": 0, "raygen": 1, "extrays": 2, "shadowrays": 3, "logic": 4, "materials": 5, "end_iteration": 6, "postprocess": 7,
              "mk_reset": 8, "mk_raygen": 9, "mk_next_vertex": 10, "mk_sample_bsdf": 11, "mk_splat": 12, "logic_fused": 13, "gather": 14}


class _PinnedBlock:
    """Owner of one flx_host_alloc block; numpy arrays made over it keep it alive through their .base chain."""

    def __init__(self, lib, nbytes):
        self._lib, self.ptr, self.nbytes = lib, C.c_void_p(), int(nbytes)
        rc = lib.flx_host_alloc(C.byref(self.ptr), self.nbytes)
        if rc != 0:
            msg = lib.flx_last_error(None)
            raise FluctusError("flx_host_alloc(%d) failed (%d): %s" % (nbytes, rc, msg.decode() if msg else "?"))
        self.buffer = (C.c_ubyte * self.nbytes).from_address(self.ptr.value)
        self.buffer._owner = self

    def __del__(self):
        try:
            if self.ptr.value:
                self._lib.flx_host_free(self.ptr)
                self.ptr = C.c_void_p()
        except Exception:
            pass


def pinned_empty(shape, dtype):
    """A numpy array in page-locked host memory (flx_host_alloc): copies between it and the device are DMA transfers at link
    speed.  The reference's analogue is the GL pixel-buffer object the picture is rendered into (src/clcontext.cpp:326-384)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    block = _PinnedBlock(_lib.load(), max(n * dtype.itemsize, 1))
    return np.frombuffer(block.buffer, dtype=dtype, count=n).reshape(shape)


def pinned_copy(a):
    out = pinned_empty(a.shape, a.dtype)
    out[...] = a
    return out


class CLContext:
    def __init__(self, num_tasks=1 << 20, device=0):  # wfBufferSize default 1<<20 (src/settings.cpp:20)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        rc = self._lib.flx_create(int(device), int(num_tasks), C.byref(self._h))
        if rc != 0:
            msg = self._lib.flx_last_error(None)
            self._h = C.c_void_p()
            raise FluctusError("flx_create failed (%d): %s" % (rc, msg.decode() if msg else "?"))
        self.NUM_TASKS = int(num_tasks)
        self._keep = []

    # ---- plumbing
    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.flx_last_error(self._h)
            raise FluctusError("%s failed (%d): %s" % (what, rc, msg.decode() if msg else "?"))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._lib.flx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @staticmethod
    def _ptr(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None and a.size else None

    # ---- setup (clcontext.hpp:62-79)
    def uploadSceneData(self, scene):
        """reference: CLContext::uploadSceneData(BVH*, Scene*) -- here the already-built arrays (fluctus_b200.SceneData)."""
        self._check(self._lib.flx_upload_scene(self._h, self._ptr(scene.tris), len(scene.tris), self._ptr(scene.indices), len(scene.indices),
                                               self._ptr(scene.nodes), len(scene.nodes), self._ptr(scene.materials), len(scene.materials),
                                               self._ptr(scene.tex_desc), len(scene.tex_desc), self._ptr(scene.tex_data), scene.tex_data.nbytes),
                    "uploadSceneData")

    def buildBVH(self, tris, max_leaf=8, quality="fast"):
        """GPU hierarchy build (flx_build_bvh): what `new SBVH(&tris, ...)` gives the reference (src/scene.cpp:574-590) -- the
        Node[] / index arrays in the reference's format -- in milliseconds.  quality: "fast" (LBVH), "ploc" (locally-ordered clustering:
        better trees, a few ms) or "ploc_opt" (ploc + the parallel-reinsertion post-pass: trees on a par with the reference's SBVH, tens of
        ms).  Returns (nodes, indices, device_ms)."""
        from .structs import NODE_DTYPE
        n = len(tris)
        nodes = np.zeros(max(2 * n - 1, 1), NODE_DTYPE)
        indices = np.zeros(n, np.uint32)
        n_nodes, ms = C.c_uint32(), C.c_float()
        self._check(self._lib.flx_build_bvh(self._h, self._ptr(tris), n, int(max_leaf), {"fast": 0, "ploc": 1, "ploc_opt": 2}[quality], self._ptr(nodes), len(nodes), C.byref(n_nodes), self._ptr(indices),
                                            C.byref(ms)), "buildBVH")
        return nodes[:n_nodes.value].copy(), indices, ms.value

    def createEnvMap(self, env):
        self._check(self._lib.flx_upload_envmap(self._h, self._ptr(env.rgb), env.width, env.height, self._ptr(env.prob), self._ptr(env.alias),
                                                self._ptr(env.pdf)), "createEnvMap")

    def setupPixelStorage(self, width, height):
        self._check(self._lib.flx_resize(self._h, int(width), int(height)), "setupPixelStorage")

    def updateParams(self, params):
        assert isinstance(params, RenderParams)
        self._check(self._lib.flx_update_params(self._h, C.byref(params)), "updateParams")

    def recompileKernels(self, setArgs=False):
        """reference: clcontext.cpp:852-874. Specialisations are chosen from the params at launch; nothing to rebuild."""

    # ---- the wavefront stages (clcontext.hpp:43-48). `params` is accepted and ignored, like in the reference,
    # where the kernels read the device copy written by updateParams.
    def enqueueWfResetKernel(self, params=None):
        self._check(self._lib.flx_enqueue_reset(self._h), "enqueueWfResetKernel")

    def enqueueWfRaygenKernel(self, params=None):
        self._check(self._lib.flx_enqueue_raygen(self._h), "enqueueWfRaygenKernel")

    def enqueueWfExtRayKernel(self, params=None):
        self._check(self._lib.flx_enqueue_extrays(self._h), "enqueueWfExtRayKernel")

    def enqueueWfShadowRayKernel(self, params=None):
        self._check(self._lib.flx_enqueue_shadowrays(self._h), "enqueueWfShadowRayKernel")

    def enqueueWfLogicKernel(self, params=None, firstIteration=False):
        self._check(self._lib.flx_enqueue_logic(self._h, 1 if firstIteration else 0), "enqueueWfLogicKernel")

    def enqueueWfMaterialKernels(self, params=None):
        self._check(self._lib.flx_enqueue_materials(self._h), "enqueueWfMaterialKernels")

    # ---- the microkernel integrator (clcontext.hpp:35-40; clcontext.cpp:709-750): one path per pixel, a phase per path
    def enqueueResetKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_reset(self._h), "enqueueResetKernel")

    def enqueueRayGenKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_raygen(self._h), "enqueueRayGenKernel")

    def enqueueNextVertexKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_next_vertex(self._h), "enqueueNextVertexKernel")

    def enqueueBsdfSampleKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_sample_bsdf(self._h), "enqueueBsdfSampleKernel")

    def enqueueSplatKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_splat(self._h), "enqueueSplatKernel")

    def enqueueSplatPreviewKernel(self, params=None):
        self._check(self._lib.flx_enqueue_mk_splat_preview(self._h), "enqueueSplatPreviewKernel")

    def renderSingleLoop(self, spp):
        """The sample loop of Tracer::renderSingle (src/tracer.cpp:124-150) spp times without host round trips (flx_render_single)."""
        self._check(self._lib.flx_render_single(self._h, int(spp)), "renderSingleLoop")

    def enqueuePostprocessKernel(self, params=None):
        """reference: clcontext.hpp:41 -- normalise / exposure / tone map / gamma into the preview buffer."""
        self._check(self._lib.flx_enqueue_postprocess(self._h), "enqueuePostprocessKernel")

    def setDenoiser(self, enabled):
        """reference: Tracer::useDenoiser (kernels rebuilt with -DUSE_OPTIX_DENOISER): accumulate the normal / albedo feature buffers."""
        self._check(self._lib.flx_set_denoiser(self._h, 1 if enabled else 0), "setDenoiser")

    def readDenoiserAOV(self, which, processed=False):
        """which: "normal" | "albedo"; processed: the display pass's output instead of the raw accumulator."""
        n = self.tilePixels()
        out = np.empty((n, 4), np.float32)
        self._check(self._lib.flx_read_denoiser_aov(self._h, {"normal": 0, "albedo": 1}[which], 1 if processed else 0, self._ptr(out), n), "readDenoiserAOV")
        return out

    def readPreview(self):
        n = self.tilePixels()
        out = np.empty((n, 4), np.float32)
        self._check(self._lib.flx_read_preview(self._h, self._ptr(out), n), "readPreview")
        return out

    # ---- queue bookkeeping (clcontext.hpp:53-57, 71)
    def enqueueClearWfQueues(self):
        self._check(self._lib.flx_enqueue_clear_queues(self._h), "enqueueClearWfQueues")

    def enqueueGetCounters(self, cnt):
        """cnt: a QueueCounters instance; filled in when finishQueue returns (reference: non-blocking read, clcontext.cpp:668-671)."""
        self._keep.append(cnt)
        self._check(self._lib.flx_enqueue_get_counters(self._h, C.byref(cnt)), "enqueueGetCounters")

    def finishQueue(self):
        self._check(self._lib.flx_finish(self._h), "finishQueue")
        self._keep.clear()

    def updatePixelIndex(self, numPixels, numNewPaths):
        self._check(self._lib.flx_update_pixel_index(self._h, int(numPixels), int(numNewPaths)), "updatePixelIndex")

    def resetPixelIndex(self):
        self._check(self._lib.flx_reset_pixel_index(self._h), "resetPixelIndex")

    def getNumTasks(self):
        return int(self._lib.flx_num_tasks(self._h))

    # ---- fused loop + statistics (new; see fluctus_b200.h)
    def render(self, iterations):
        self._check(self._lib.flx_render(self._h, int(iterations)), "render")

    def renderTimed(self, iterations):
        """flx_render bracketed by CUDA events on the context's stream; returns device milliseconds."""
        ms = C.c_float()
        self._check(self._lib.flx_render_timed(self._h, int(iterations), C.byref(ms)), "renderTimed")
        return ms.value

    def timerBegin(self):
        self._check(self._lib.flx_timer_begin(self._h), "timerBegin")

    def timerEnd(self):
        ms = C.c_float()
        self._check(self._lib.flx_timer_end(self._h, C.byref(ms)), "timerEnd")
        return ms.value

    TUNING = {"trace_variant": 0, "fetch_threshold": 1, "trace_blocks_per_sm": 2, "top_nodes": 3, "inner_min": 4, "logic_min_blocks": 5, "fetch_chunk": 6, "overlap_trace": 7, "postprocess_in_loop": 10, "smem_stack": 11, "max_l1": 12, "fuse_stages": 13, "prefetch_children": 15, "repack_on_host": 16, "overlap_postprocess": 17, "dirty_postprocess": 18, "l2_persist": 19, "fused_min_blocks": 14, "ext_min_blocks": 8, "shadow_min_blocks": 9, "inner_bias": 20, "gather_priority": 21, "bvh_tri_cost": 22, "gather_direct": 23, "logic_tile": 24, "shadow_left_first": 25, "bvh_reinsert": 26, "material_mask": 27, "bvh_depth_limit": 28}

    def setTuning(self, **kv):
        for k, v in kv.items():
            self._check(self._lib.flx_set_tuning(self._h, self.TUNING[k], int(v)), "setTuning(%s)" % k)

    def setCounting(self, enabled):
        self._check(self._lib.flx_set_counting(self._h, 1 if enabled else 0), "setCounting")

    def getTraceCounts(self):
        """{'ext': {...}, 'shadow': {...}} totals of nodes/boxes/tris/updates/rays since setCounting(True)."""
        a = (C.c_uint64 * 5)()
        b = (C.c_uint64 * 5)()
        self._check(self._lib.flx_get_trace_counts(self._h, a, b), "getTraceCounts")
        keys = ("nodes", "boxes", "tris", "updates", "rays")
        return {"ext": dict(zip(keys, map(int, a))), "shadow": dict(zip(keys, map(int, b)))}

    def resetStats(self):
        self._check(self._lib.flx_reset_stats(self._h), "resetStats")

    def getStats(self):
        s = RenderStats64()
        self._check(self._lib.flx_get_stats(self._h, C.byref(s)), "getStats")
        return s

    def setProfiling(self, enabled):
        self._check(self._lib.flx_set_profiling(self._h, 1 if enabled else 0), "setProfiling")

    def checkTracingPerf(self):
        """reference: clcontext.cpp:673-701 -- per-kernel device time; returns {name: (total_ms, launches)}."""
        out = {}
        for name, kid in KERNEL_IDS.items():
            ms, n = C.c_float(), C.c_uint32()
            self._check(self._lib.flx_get_kernel_ms(self._h, kid, C.byref(ms), C.byref(n)), "checkTracingPerf")
            out[name] = (ms.value, n.value)
        return out

    # ---- read-back
    def tilePixels(self):
        return int(self._lib.flx_tile_pixels(self._h))

    def readPixels(self, out=None):
        """The accumulator (RGB sums, sample count) of this context's pixels.  `out`: a caller's (tilePixels, 4) float32 array to
        fill -- e.g. one from pinned_empty(), which makes the copy a DMA transfer."""
        n = self.tilePixels()
        if out is None:
            out = np.empty((n, 4), np.float32)
        assert out.dtype == np.float32 and out.size == n * 4 and out.flags.c_contiguous
        self._check(self._lib.flx_read_pixels(self._h, self._ptr(out), n), "readPixels")
        return out

    def saveImage(self, filename, params=None):
        """reference: CLContext::saveImage (clcontext.cpp:386-465) -- '*.hdr': linear radiance as Radiance RGBE, otherwise the
        post-processed preview as 8-bit PNG."""
        self._check(self._lib.flx_save_image(self._h, str(filename).encode()), "saveImage")

    def saveCheckpoint(self, path):
        """Everything an interrupted render needs to continue (path state, queues, counters, pixel index, statistics, accumulator)."""
        self._check(self._lib.flx_checkpoint_save(self._h, str(path).encode()), "saveCheckpoint")

    def loadCheckpoint(self, path):
        self._check(self._lib.flx_checkpoint_load(self._h, str(path).encode()), "loadCheckpoint")

    def readTraversalLayout(self):
        """(tnodes (n, 16) float32, ttris (m, 16) float32, rootRef): the uploaded hierarchy in the traversal layout (diagnostic)."""
        nn, nt, root = C.c_uint32(), C.c_uint32(), C.c_int32()
        self._check(self._lib.flx_read_traversal_layout(self._h, None, C.byref(nn), None, C.byref(nt), C.byref(root)), "readTraversalLayout")
        tn, tt = np.zeros((nn.value, 16), np.float32), np.zeros((nt.value, 16), np.float32)
        self._check(self._lib.flx_read_traversal_layout(self._h, self._ptr(tn) if tn.size else tn.ctypes.data_as(C.c_void_p), C.byref(nn),
                                                        self._ptr(tt) if tt.size else tt.ctypes.data_as(C.c_void_p), C.byref(nt), C.byref(root)), "readTraversalLayout")
        return tn, tt, root.value

    def readTasks(self):
        out = np.empty((64, self.NUM_TASKS), np.uint32)
        self._check(self._lib.flx_read_tasks(self._h, self._ptr(out)), "readTasks")
        return out

    def writeTasks(self, slots):
        slots = np.ascontiguousarray(slots, np.uint32)
        assert slots.shape == (64, self.NUM_TASKS)
        self._check(self._lib.flx_write_tasks(self._h, self._ptr(slots)), "writeTasks")

    def readQueue(self, name, n=None):
        out = np.empty(self.NUM_TASKS if n is None else int(n), np.uint32)
        if out.size == 0:
            return out
        self._check(self._lib.flx_read_queue(self._h, QUEUE_NAMES.index(name), self._ptr(out), len(out)), "readQueue")
        return out

    def writeQueue(self, name, entries):
        entries = np.ascontiguousarray(entries, np.uint32)
        if entries.size == 0:
            return
        self._check(self._lib.flx_write_queue(self._h, QUEUE_NAMES.index(name), self._ptr(entries), len(entries)), "writeQueue")

    def writeCounters(self, cnt):
        self._check(self._lib.flx_write_counters(self._h, C.byref(cnt)), "writeCounters")

    def readCounters(self):
        cnt = QueueCounters()
        self.enqueueGetCounters(cnt)
        self.finishQueue()
        return cnt

    def deviceBytes(self):
        return int(self._lib.flx_device_bytes(self._h))

    # ---- multi-GPU
    def setTile(self, part, n_parts, stripe_rows=8):
        self._check(self._lib.flx_set_tile(self._h, int(part), int(n_parts), int(stripe_rows)), "setTile")

    def commUniqueId(self):
        buf = (C.c_char * 128)()
        rc = self._lib.flx_comm_unique_id(buf)
        if rc != 0:
            raise FluctusError("flx_comm_unique_id failed (%d): %s" % (rc, (self._lib.flx_last_error(None) or b"?").decode()))
        return bytes(buf)

    def commInit(self, unique_id, rank, nranks):
        buf = C.create_string_buffer(bytes(unique_id), 128)
        self._check(self._lib.flx_comm_init(self._h, buf, int(rank), int(nranks)), "commInit")

    def gatherPixels(self, root=0, out=None):
        self._check(self._lib.flx_gather_pixels(self._h, int(root), self._ptr(out) if out is not None else None), "gatherPixels")

    def readGathered(self, width, height, preview=False):
        """Root only, after gatherPixels: the gathered full image -- accumulators, or (preview) the display pass over them."""
        out = np.empty((int(width) * int(height), 4), np.float32)
        self._check(self._lib.flx_read_gathered(self._h, 1 if preview else 0, self._ptr(out), len(out)), "readGathered")
        return out
