"""Tracer -- the part of the reference's Tracer that drives the wavefront path, headless.

Replays, call for call, the order in which the reference drives CLContext:
  * start()     = the `iteration == 0` prologue of Tracer::update       (reference: src/tracer.cpp:236-240)
  * iterate()   = one steady-state iteration of Tracer::runBenchmark    (reference: src/tracer.cpp:433-439, 447 [post-process], 455-465)
  * render(n)   = n such iterations without host round trips (flx_render)
  * renderSingle(spp) / updateMicrokernel() = the same two loops with the reference's other integrator, the microkernel
                  path tracer                                            (reference: src/tracer.cpp:95-169, 267-299)
It works with any object that has CLContext's method set -- the CUDA context (fluctus_b200.CLContext) and the oracle
contexts in oracle/ alike -- which is how the parity tests run the same loop on both.
"""
from .structs import QueueCounters


class Tracer:
    def __init__(self, clctx, params):
        self.clctx = clctx
        self.params = params
        self.iteration = 0
        self.stats = dict(primaryRays=0, extensionRays=0, shadowRays=0, samples=0)

    def start(self):
        c, p = self.clctx, self.params
        c.updateParams(p)
        c.resetPixelIndex()
        c.enqueueWfResetKernel(p)   # puts all paths in the raygen queue
        c.enqueueWfRaygenKernel(p)
        c.enqueueWfExtRayKernel(p)
        c.enqueueClearWfQueues()
        c.finishQueue()
        self.iteration = 0

    def iterate(self):
        c, p = self.clctx, self.params
        cnt = QueueCounters()
        c.enqueueWfLogicKernel(p, False)
        c.enqueueWfRaygenKernel(p)
        c.enqueueWfMaterialKernels(p)
        c.enqueueGetCounters(cnt)   # the following kernels do not grow the queues
        c.enqueueWfExtRayKernel(p)
        c.enqueueWfShadowRayKernel(p)
        c.enqueueClearWfQueues()
        c.enqueuePostprocessKernel(p)  # src/tracer.cpp:302, 447: the display pass is part of every loop iteration
        c.finishQueue()
        self.stats["extensionRays"] += cnt.extensionQueue
        self.stats["shadowRays"] += cnt.shadowQueue
        self.stats["primaryRays"] += cnt.raygenQueue
        self.stats["samples"] += cnt.raygenQueue
        c.updatePixelIndex(self._num_pixels(), cnt.raygenQueue)
        self.iteration += 1
        return cnt

    def update(self):
        """One call of the interactive loop, Tracer::update() wavefront branch (src/tracer.cpp:222-266, 301-308, 333-340):
        on `iteration == 0` the accumulation restarts with a preview -- maxBounces clamped to 2, the prologue, then THREE
        logic/raygen/material/trace rounds with `firstIteration` set, and only the LAST round's counters advance the
        pixel index -- otherwise one normal round."""
        c, p = self.clctx, self.params
        cnt = QueueCounters()
        n_rounds = 1
        max_bounces = p.maxBounces
        first = self.iteration == 0
        if first:
            p.maxBounces = min(2, max_bounces)
            c.updateParams(p)
            n_rounds = 3
            c.resetPixelIndex()
            c.enqueueWfResetKernel(p)
            c.enqueueWfRaygenKernel(p)
            c.enqueueWfExtRayKernel(p)
            c.enqueueClearWfQueues()
        for _ in range(n_rounds):
            cnt = QueueCounters()
            c.enqueueWfLogicKernel(p, first)
            c.enqueueWfRaygenKernel(p)
            c.enqueueWfMaterialKernels(p)
            c.enqueueGetCounters(cnt)
            c.enqueueWfExtRayKernel(p)
            c.enqueueWfShadowRayKernel(p)
            c.enqueueClearWfQueues()
        if first:
            p.maxBounces = max_bounces
            c.updateParams(p)
        c.enqueuePostprocessKernel(p)
        c.finishQueue()
        c.updatePixelIndex(self._num_pixels(), cnt.raygenQueue)
        self.stats["extensionRays"] += cnt.extensionQueue
        self.stats["shadowRays"] += cnt.shadowQueue
        self.stats["primaryRays"] += cnt.raygenQueue
        self.stats["samples"] += cnt.raygenQueue if self.iteration > 0 else 0
        self.iteration += 1
        return cnt

    # ---- the microkernel integrator (the reference's default, `useWavefront = false`)
    def renderSingle(self, spp, fused=False):
        """Tracer::renderSingle (src/tracer.cpp:95-169): a final frame with exactly `spp` samples in every pixel -- which only
        the microkernel integrator guarantees, so the reference switches to it here (tracer.cpp:99-101) and turns Russian
        roulette off.  Per sample: camera rays, (maxBounces + 1) x (nextVertex, sampleBsdf), splat, display pass, finish.
        `fused=True` runs the same loop inside the library without the per-sample finishQueue (flx_render_single)."""
        c, p = self.clctx, self.params
        if p.useRoulette:
            p.useRoulette = 0
        c.updateParams(p)
        c.enqueueResetKernel(p)
        if fused and hasattr(c, "renderSingleLoop"):
            c.renderSingleLoop(spp)
            c.finishQueue()
            return
        for _ in range(spp):
            c.enqueueRayGenKernel(p)
            for _bounce in range(p.maxBounces + 1):
                c.enqueueNextVertexKernel(p)
                c.enqueueBsdfSampleKernel(p)
            c.enqueueSplatKernel(p)
            c.enqueuePostprocessKernel(p)
            c.finishQueue()

    def updateMicrokernel(self):
        """One call of the interactive loop with the microkernel integrator, Tracer::update() `else` branch
        (src/tracer.cpp:267-299): the first call after a change resets and shows a two-segment preview, later calls add
        one path segment per call."""
        c, p = self.clctx, self.params
        if self.iteration == 0:
            c.updateParams(p)
            c.enqueueResetKernel(p)
            c.enqueueRayGenKernel(p)
            c.enqueueNextVertexKernel(p)
            c.enqueueBsdfSampleKernel(p)
            c.enqueueNextVertexKernel(p)
            c.enqueueBsdfSampleKernel(p)
            c.enqueueSplatPreviewKernel(p)
        else:
            c.enqueueRayGenKernel(p)
            c.enqueueNextVertexKernel(p)
            c.enqueueBsdfSampleKernel(p)
            c.enqueueSplatKernel(p)
        c.enqueuePostprocessKernel(p)
        c.finishQueue()
        self.iteration += 1

    def _num_pixels(self):
        tp = getattr(self.clctx, "tilePixels", None)
        return tp() if tp else self.params.width * self.params.height

    def render(self, iterations):
        """Fused loop (device-side bookkeeping); falls back to iterate() for contexts without it (the oracles)."""
        if hasattr(self.clctx, "render"):
            self.clctx.render(iterations)
            self.iteration += iterations
        else:
            for _ in range(iterations):
                self.iterate()
