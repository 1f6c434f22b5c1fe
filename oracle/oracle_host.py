"""oracle_host.py -- TEST INFRASTRUCTURE. CPU contexts with CLContext's method set, for parity checks only.

Two oracles share this driver (only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
leg may import this module; the product package never does):

  RefContext   runs the REFERENCE'S OWN kernel sources compiled for the host (oracle/_ref/libfluctus_ref.so, built by
               oracle/build_ref.py from /root/reference/src/wf_*.cl).  Serial = deterministic ground truth; `parallel=True`
               = the OpenMP build used as the CPU baseline.
  PortContext  runs the plain-C restatement in oracle/wf_oracle.c (oracle/liboracle.so).

Both keep every buffer the reference keeps in cl::Buffers (src/clcontext.hpp:167-210) as numpy arrays and enqueue
"kernels" by looping the NDRange the reference launches (src/clcontext.cpp:765-848).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_LIB = os.path.join(HERE, "_ref", "libfluctus_ref.so")
PORT_LIB = os.path.join(HERE, "liboracle.so")


class RefBufs(C.Structure):  # oracle/ref_shim/ref_abi.h
    _fields_ = [("tasks", C.c_void_p), ("pixels", C.c_void_p), ("denoiserAlbedo", C.c_void_p), ("denoiserNormal", C.c_void_p),
                ("queueLens", C.c_void_p), ("raygenQueue", C.c_void_p), ("extensionQueue", C.c_void_p), ("shadowQueue", C.c_void_p),
                ("diffuseQueue", C.c_void_p), ("glossyQueue", C.c_void_p), ("ggxReflQueue", C.c_void_p), ("ggxRefrQueue", C.c_void_p),
                ("deltaQueue", C.c_void_p), ("tris", C.c_void_p), ("nodes", C.c_void_p), ("indices", C.c_void_p), ("envRGBA", C.c_void_p),
                ("envW", C.c_int32), ("envH", C.c_int32), ("probTable", C.c_void_p), ("aliasTable", C.c_void_p), ("pdfTable", C.c_void_p),
                ("materials", C.c_void_p), ("texData", C.c_void_p), ("textures", C.c_void_p), ("params", C.c_void_p),
                ("currPixelIdx", C.c_void_p), ("numTasks", C.c_uint32), ("firstIteration", C.c_uint32), ("pixelsPreview", C.c_void_p),
                ("stats", C.c_void_p), ("denoiserAlbedoGL", C.c_void_p), ("denoiserNormalGL", C.c_void_p)]


QUEUE_FIELDS = ("raygenQueue", "extensionQueue", "shadowQueue", "diffuseQueue", "glossyQueue", "ggxReflQueue", "ggxRefrQueue", "deltaQueue")
QUEUE_NAMES = ("raygen", "extension", "shadow", "diffuse", "glossy", "ggxRefl", "ggxRefr", "delta")


def ref_available():
    return os.path.exists(REF_LIB)


def port_available():
    return os.path.exists(PORT_LIB)


class _CpuContext:
    """Shared implementation; subclasses pick the library and symbol prefix."""

    LIB = None
    PREFIX = None

    def __init__(self, num_tasks, parallel=False, parallel_trace=False, variant=""):
        """variant: one of the extra builds of oracle/build_ref.py -- "dn_" (USE_OPTIX_DENOISER: feature buffers), "bs_" (USE_BITSTACK
        traversal), "aos_" (no USE_SOA: array-of-structures path state), "lm_" (C-library math instead of include/flx_math.h).  A
        kernel the variant does not rebuild comes from the base build.
        parallel: every kernel through the OpenMP build (CPU baseline; queue order then depends on thread timing).
        parallel_trace: only the two traversal kernels through the OpenMP build -- they push to no queue and every work-item
        writes its own path's slots only, so the result is the serial one; what stays serial (reset, raygen, logic, materials)
        is what decides queue order.  This is the deterministic oracle for the full-size parity tests."""
        if not os.path.exists(self.LIB):
            raise FileNotFoundError(self.LIB)
        self.lib = C.CDLL(self.LIB)
        self.prefix = self.PREFIX + ("par_" if parallel else "")
        self.trace_prefix = self.PREFIX + ("par_" if (parallel or parallel_trace) else "")
        self.variant = variant
        self.aos = variant == "aos_"
        self.NUM_TASKS = int(num_tasks)
        n = self.NUM_TASKS
        # path state: structure of arrays [slot][path] -- or, in the AoS build, N structs of 64 words (src/geom.h:199-236)
        self.tasks = np.zeros((n, 64) if self.aos else (64, n), np.uint32)
        self.queues = {q: np.zeros(n, np.uint32) for q in QUEUE_FIELDS}
        self.counters = np.zeros(8, np.uint32)
        self.currPixelIdx = np.zeros(1, np.uint32)
        self.render_stats = np.zeros(4, np.uint32)  # RenderStats {primaryRays, extensionRays, shadowRays, samples}, geom.h:254-260
        self.pixelIdx = 0
        self.params_buf = np.zeros(240, np.uint8)
        self.env_rgba = np.zeros(4, np.float32)
        self.env_w = self.env_h = 1
        self.prob = np.zeros(1, np.float32)
        self.alias = np.zeros(1, np.int32)
        self.pdf = np.zeros(1, np.float32)
        self.scene = None
        self.pixels = self.albedo = self.normal = None
        self._pending = []
        self._separate = False
        self.width = self.height = 0

    # ---- setup
    def uploadSceneData(self, scene):
        self.scene = scene
        self.tex_desc = scene.tex_desc if len(scene.tex_desc) else np.zeros(12, np.uint8)
        self.tex_data = scene.tex_data if len(scene.tex_data) else np.zeros(4, np.uint8)

    def createEnvMap(self, env):
        n = env.width * env.height
        rgba = np.ones((n, 4), np.float32)
        rgba[:, :3] = env.rgb.reshape(n, 3)
        self.env_rgba, self.env_w, self.env_h = np.ascontiguousarray(rgba), env.width, env.height
        self.prob, self.alias, self.pdf = env.prob, env.alias, env.pdf

    def setupPixelStorage(self, width, height):
        self.width, self.height = int(width), int(height)
        self.pixels = np.zeros((self.width * self.height, 4), np.float32)
        self.albedo = np.zeros_like(self.pixels)
        self.normal = np.zeros_like(self.pixels)
        self.preview = np.zeros_like(self.pixels)
        self.albedo_out = np.zeros_like(self.pixels)
        self.normal_out = np.zeros_like(self.pixels)

    def updateParams(self, params):
        C.memmove(self.params_buf.ctypes.data, C.byref(params), 240)
        self._separate = bool(params.wfSeparateQueues)

    def recompileKernels(self, setArgs=False):
        pass

    def _bufs(self, first=0):
        b = RefBufs()
        p = lambda a: a.ctypes.data
        b.tasks, b.pixels, b.denoiserAlbedo, b.denoiserNormal = p(self.tasks), p(self.pixels), p(self.albedo), p(self.normal)
        b.queueLens = p(self.counters)
        for q in QUEUE_FIELDS:
            setattr(b, q, p(self.queues[q]))
        s = self.scene
        b.tris, b.nodes, b.indices, b.materials = p(s.tris), p(s.nodes), p(s.indices), p(s.materials)
        b.texData, b.textures = p(self.tex_data), p(self.tex_desc)
        b.envRGBA, b.envW, b.envH = p(self.env_rgba), self.env_w, self.env_h
        b.probTable, b.aliasTable, b.pdfTable = p(self.prob), p(self.alias), p(self.pdf)
        b.params, b.currPixelIdx = p(self.params_buf), p(self.currPixelIdx)
        b.numTasks, b.firstIteration = self.NUM_TASKS, first
        b.pixelsPreview = p(self.preview)
        b.stats = p(self.render_stats)
        b.denoiserAlbedoGL, b.denoiserNormalGL = p(self.albedo_out), p(self.normal_out)
        return b

    def _run(self, name, n, first=0):
        base = self.trace_prefix if name in ("ext", "shadow") else self.prefix
        try:
            fn = getattr(self.lib, base + self.variant + name)
        except AttributeError:
            if self.aos:
                raise  # every kernel must agree on the path-state layout
            fn = getattr(self.lib, base + name)
        fn.argtypes = [C.POINTER(RefBufs), C.c_size_t, C.c_size_t]
        fn.restype = None
        b = self._bufs(first)
        fn(C.byref(b), 0, int(n))

    # ---- stages: NDRange sizes of src/clcontext.cpp:765-848
    def enqueueWfResetKernel(self, params=None):
        self._run("reset", max(self.NUM_TASKS, self.width * self.height))

    def enqueueWfRaygenKernel(self, params=None):
        self._run("raygen", self.NUM_TASKS)

    def enqueueWfExtRayKernel(self, params=None):
        self._run("ext", self.NUM_TASKS)

    def enqueueWfShadowRayKernel(self, params=None):
        self._run("shadow", self.NUM_TASKS)

    def enqueueWfLogicKernel(self, params=None, firstIteration=False):
        n = ((self.NUM_TASKS - 1) // 32 + 1) * 32
        self._run("logic_separate" if self._separate else "logic_single", n, 1 if firstIteration else 0)

    def enqueueWfMaterialKernels(self, params=None):
        if self._separate:
            for k in ("mat_diffuse", "mat_glossy", "mat_ggx_refl", "mat_ggx_refr", "mat_delta"):
                self._run(k, self.NUM_TASKS)
        else:
            self._run("mat_all", self.NUM_TASKS)

    # ---- microkernel integrator: NDRange sizes of src/clcontext.cpp:709-750
    def enqueueResetKernel(self, params=None):
        self._run("mk_reset", self.width * self.height)

    def enqueueRayGenKernel(self, params=None):
        self._run("mk_raygen", self.NUM_TASKS)

    def enqueueNextVertexKernel(self, params=None):
        self._run("mk_next_vertex", self.NUM_TASKS)

    def enqueueBsdfSampleKernel(self, params=None):
        self._run("mk_sample_bsdf", self.NUM_TASKS)

    def enqueueSplatKernel(self, params=None):
        self._run("mk_splat", self.width * self.height)

    def enqueueSplatPreviewKernel(self, params=None):
        self._run("mk_splat_preview", self.width * self.height)

    def resetStats(self):
        self.render_stats[:] = 0

    def getStats(self):
        from fluctus_b200.structs import RenderStats64
        s = RenderStats64()
        s.primaryRays, s.extensionRays, s.shadowRays, s.samples = (int(v) for v in self.render_stats)
        return s

    def enqueuePostprocessKernel(self, params=None):
        self._run("postprocess", self.width * self.height)  # NDRange(width*height), src/clcontext.cpp:758

    def readPreview(self):
        return self.preview.copy()

    # ---- bookkeeping
    def enqueueClearWfQueues(self):
        self.counters[:] = 0

    def enqueueGetCounters(self, cnt):
        for i, (name, _) in enumerate(cnt._fields_):
            setattr(cnt, name, int(self.counters[i]))

    def finishQueue(self):
        pass

    def updatePixelIndex(self, numPixels, numNewPaths):
        self.pixelIdx = (self.pixelIdx + int(numNewPaths)) % int(numPixels)
        self.currPixelIdx[0] = self.pixelIdx

    def resetPixelIndex(self):
        self.pixelIdx = 0
        self.currPixelIdx[0] = 0

    def getNumTasks(self):
        return self.NUM_TASKS

    # ---- read-back, same names as fluctus_b200.CLContext
    def tilePixels(self):
        return self.width * self.height

    def readPixels(self):
        return self.pixels.copy()

    def readTasks(self):
        return np.ascontiguousarray(self.tasks.T) if self.aos else self.tasks.copy()

    def writeTasks(self, slots):
        self.tasks[...] = slots.T if self.aos else slots

    def setDenoiser(self, enabled):
        assert bool(enabled) == (self.variant == "dn_"), "the feature buffers are a build option of the reference: RefContext(n, variant='dn_')"

    def readDenoiserAOV(self, which, processed=False):
        src = {("normal", False): self.normal, ("albedo", False): self.albedo, ("normal", True): self.normal_out, ("albedo", True): self.albedo_out}
        return src[(which, bool(processed))].copy()

    def readQueue(self, name, n=None):
        q = self.queues[QUEUE_FIELDS[QUEUE_NAMES.index(name)]]
        return q.copy() if n is None else q[:n].copy()

    def writeQueue(self, name, entries):
        q = self.queues[QUEUE_FIELDS[QUEUE_NAMES.index(name)]]
        q[:len(entries)] = entries

    def writeCounters(self, cnt):
        for i, (name, _) in enumerate(cnt._fields_):
            self.counters[i] = getattr(cnt, name)

    def readCounters(self):
        from fluctus_b200.structs import QueueCounters
        cnt = QueueCounters()
        self.enqueueGetCounters(cnt)
        return cnt

    def close(self):
        pass


class RefContext(_CpuContext):
    LIB = REF_LIB
    PREFIX = "ref_"


class PortContext(_CpuContext):
    LIB = PORT_LIB
    PREFIX = "port_"


def build_ploc(tris, max_leaf=8, tri_cost=1.0, reinsert=0, depth_limit=62):
    """CPU restatement of the FLX_BVH_PLOC builder (locally-ordered clustering) -> (nodes, indices); reinsert = iterations of
    the parallel-reinsertion post-pass (FLX_BVH_PLOC_OPT: FLX_TUNE_BVH_REINSERT iterations)."""
    return build_lbvh(tris, max_leaf, fn="port_build_ploc", tri_cost=tri_cost, reinsert=reinsert, depth_limit=depth_limit)


def build_lbvh(tris, max_leaf=8, fn="port_build_lbvh", tri_cost=1.0, reinsert=0, depth_limit=62):
    """CPU restatement (oracle/bvh_oracle.c) of the GPU hierarchy builder flx_build_bvh -> (nodes, indices) in the
    reference's Node[] / index-list format.  tri_cost = FLX_TUNE_BVH_TRI_COST / 100."""
    from fluctus_b200.structs import NODE_DTYPE
    lib = C.CDLL(PORT_LIB)
    lib.port_set_tri_cost.argtypes = [C.c_float]
    lib.port_set_tri_cost(float(tri_cost))
    lib.port_set_reinsert.argtypes = [C.c_int]
    lib.port_set_reinsert(int(reinsert))
    lib.port_set_depth_limit.argtypes = [C.c_int]
    lib.port_set_depth_limit(int(depth_limit))
    n = len(tris)
    nodes = np.zeros(max(2 * n - 1, 1), NODE_DTYPE)
    indices = np.zeros(n, np.uint32)
    n_nodes = C.c_uint32()
    f = getattr(lib, fn)
    f.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.POINTER(C.c_uint32), C.c_void_p]
    rc = f(tris.ctypes.data, n, int(max_leaf), nodes.ctypes.data, C.byref(n_nodes), indices.ctypes.data)
    lib.port_set_tri_cost(1.0)
    lib.port_set_reinsert(0)
    lib.port_set_depth_limit(62)
    if rc != 0:
        raise RuntimeError("%s failed (%d)" % (fn, rc))
    return nodes[:n_nodes.value].copy(), indices
