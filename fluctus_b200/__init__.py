"""fluctus_b200 -- B200-native wavefront path-tracing hot path behind the reference's CLContext API.

The product is the CUDA library (csrc/ -> libfluctus_b200.so, C ABI in include/fluctus_b200.h).
This package is the thin host-side mirror of the reference's CLContext / Tracer loop
(reference: src/clcontext.hpp, src/tracer.cpp:222-266, 431-470) used by the tests and bench.py.
There is no CPU fallback: importing CLContext works anywhere, creating one needs a B200.
"""
from .structs import (RenderParams, QueueCounters, RenderStats64, Camera, AreaLight, NODE_DTYPE, TRIANGLE_DTYPE, MATERIAL_DTYPE,
                      TEXDESC_DTYPE, SLOT, BXDF)
from .scene import SceneData, EnvMapData, make_params, look_at
from .clcontext import CLContext, FluctusError, pinned_empty, pinned_copy
from .tracer import Tracer

__all__ = ["CLContext", "FluctusError", "pinned_empty", "pinned_copy", "Tracer", "SceneData", "EnvMapData", "RenderParams", "QueueCounters", "RenderStats64", "Camera",
           "AreaLight", "make_params", "look_at", "NODE_DTYPE", "TRIANGLE_DTYPE", "MATERIAL_DTYPE", "TEXDESC_DTYPE", "SLOT", "BXDF"]
