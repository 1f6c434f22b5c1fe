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

    # ---- the reference's own benchmark protocol
    def runBenchmarkScene(self, scene_name, render_len=30.0, use_wavefront=True, log_every=0.5, clock=None):
        """One scene of Tracer::runBenchmark (src/tracer.cpp:372-383, 416-503): reset BOTH integrators' state and the queue
        counters (no prologue: the first logic pass finds every path at length 0 and sends it to raygen), then iterate for
        `render_len` seconds of wall time, synchronising every iteration, and log Mrays/s every `log_every` seconds as CSV
        rows `scene;time;primary;extension;shadow;total;samples` (the format plot_benchmarks.py reads).  Returns
        (rows, summary) where summary = Mrays/s over the whole run: primary, extension, shadow, samples, total."""
        import time
        clock = clock or time.perf_counter
        c, p = self.clctx, self.params
        self.iteration = 0
        c.updateParams(p)
        c.enqueueResetKernel(p)
        c.enqueueWfResetKernel(p)
        c.enqueueClearWfQueues()
        c.finishQueue()
        c.resetStats()
        rows, log = [], []
        acc = dict(primaryRays=0, extensionRays=0, shadowRays=0, samples=0)
        mk_prev = (0, 0, 0, 0)
        start = curr = last_log = clock()

        def log_stats(elapsed, delta_t):
            nonlocal acc, last_log
            s = 1e6 * delta_t
            log.append(dict(acc))
            rows.append("%s;%g;%g;%g;%g;%g;%g" % (scene_name, elapsed, acc["primaryRays"] / s, acc["extensionRays"] / s, acc["shadowRays"] / s,
                                                  (acc["primaryRays"] + acc["extensionRays"] + acc["shadowRays"]) / s, acc["samples"] / s))
            acc = dict(primaryRays=0, extensionRays=0, shadowRays=0, samples=0)
            last_log = clock()

        while curr - start < render_len:
            cnt = QueueCounters()
            if use_wavefront:
                c.enqueueWfLogicKernel(p, False)
                c.enqueueWfRaygenKernel(p)
                c.enqueueWfMaterialKernels(p)
                c.enqueueGetCounters(cnt)
                c.enqueueWfExtRayKernel(p)
                c.enqueueWfShadowRayKernel(p)
                c.enqueueClearWfQueues()
            else:
                c.enqueueRayGenKernel(p)
                c.enqueueNextVertexKernel(p)
                c.enqueueBsdfSampleKernel(p)
                c.enqueueSplatKernel(p)
            c.enqueuePostprocessKernel(p)
            c.finishQueue()
            if use_wavefront:
                acc["extensionRays"] += cnt.extensionQueue
                acc["shadowRays"] += cnt.shadowQueue
                acc["primaryRays"] += cnt.raygenQueue
                acc["samples"] += cnt.raygenQueue if self.iteration > 0 else 0
            else:  # fetchStatsAsync: the microkernels count on the device
                st = c.getStats()
                now = (int(st.primaryRays), int(st.extensionRays), int(st.shadowRays), int(st.samples))
                for k, a, b in zip(("primaryRays", "extensionRays", "shadowRays", "samples"), now, mk_prev):
                    acc[k] += a - b
                mk_prev = now
            c.updatePixelIndex(self._num_pixels(), cnt.raygenQueue)
            if curr - last_log > log_every:
                log_stats(curr - start, curr - last_log)
            self.iteration += 1
            curr = clock()
        log_stats(curr - start, max(curr - last_log, 1e-9))
        total_t = 1e6 * (curr - start)
        sums = {k: sum(entry[k] for entry in log) for k in ("primaryRays", "extensionRays", "shadowRays", "samples")}
        summary = dict(primary=sums["primaryRays"] / total_t, extension=sums["extensionRays"] / total_t, shadow=sums["shadowRays"] / total_t,
                       samples=sums["samples"] / total_t, total=(sums["primaryRays"] + sums["extensionRays"] + sums["shadowRays"]) / total_t,
                       iterations=self.iteration, seconds=curr - start)
        return rows, summary

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
