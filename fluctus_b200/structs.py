"""Byte layouts shared with the device (reference: src/geom.h; C mirror: include/fluctus_b200.h)."""
import ctypes as C

import numpy as np


class Float3(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("z", C.c_float), ("w", C.c_float)]

    def set(self, v):
        self.x, self.y, self.z = float(v[0]), float(v[1]), float(v[2])
        self.w = float(v[3]) if len(v) > 3 else 0.0

    def tolist(self):
        return [self.x, self.y, self.z]


class Float2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class AreaLight(C.Structure):  # geom.h:104-111
    _fields_ = [("right", Float3), ("up", Float3), ("N", Float3), ("pos", Float3), ("E", Float3), ("size", Float2), ("_pad", C.c_float * 2)]


class Camera(C.Structure):  # geom.h:146-155
    _fields_ = [("pos", Float3), ("dir", Float3), ("up", Float3), ("right", Float3), ("fov", C.c_float), ("apertureSize", C.c_float),
                ("focalDist", C.c_float), ("_pad", C.c_float)]


class PostProcessParams(C.Structure):
    _fields_ = [("exposure", C.c_float), ("tmOperator", C.c_uint32)]


class RenderParams(C.Structure):  # geom.h:163-180
    _fields_ = [("areaLight", AreaLight), ("camera", Camera), ("ppParams", PostProcessParams), ("width", C.c_uint32), ("height", C.c_uint32),
                ("n_tris", C.c_uint32), ("useEnvMap", C.c_uint32), ("useAreaLight", C.c_uint32), ("envMapStrength", C.c_float),
                ("maxBounces", C.c_uint32), ("sampleImpl", C.c_uint32), ("sampleExpl", C.c_uint32), ("useRoulette", C.c_uint32),
                ("wfSeparateQueues", C.c_uint32), ("worldRadius", C.c_float), ("_pad", C.c_uint32 * 2)]

    def copy(self):
        other = RenderParams()
        C.memmove(C.byref(other), C.byref(self), C.sizeof(RenderParams))
        return other


class QueueCounters(C.Structure):  # geom.h:240-252
    _fields_ = [(n, C.c_uint32) for n in ("raygenQueue", "extensionQueue", "shadowQueue", "diffuseQueue", "glossyQueue", "ggxReflQueue",
                                          "ggxRefrQueue", "deltaQueue")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class RenderStats64(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("primaryRays", "extensionRays", "shadowRays", "samples", "iterations")]


assert C.sizeof(RenderParams) == 240 and C.sizeof(AreaLight) == 96 and C.sizeof(Camera) == 80 and C.sizeof(QueueCounters) == 32

F3 = [("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("w", "<f4")]
NODE_DTYPE = np.dtype([("bmin", "<f4", 4), ("bmax", "<f4", 4), ("parent", "<i4"), ("link", "<u4"), ("nPrims", "u1"), ("_pad", "u1", 7)])  # 48 B
VERTEX_DTYPE = np.dtype([("p", "<f4", 4), ("n", "<f4", 4), ("t", "<f4", 4)])
TRIANGLE_DTYPE = np.dtype([("v0", VERTEX_DTYPE), ("v1", VERTEX_DTYPE), ("v2", VERTEX_DTYPE), ("matId", "<i4"), ("_pad", "<i4", 3)])  # 160 B
MATERIAL_DTYPE = np.dtype([("Kd", "<f4", 4), ("Ks", "<f4", 4), ("Ke", "<f4", 4), ("Ns", "<f4"), ("Ni", "<f4"), ("map_Kd", "<i4"), ("map_Ks", "<i4"),
                           ("map_N", "<i4"), ("type", "<i4"), ("_pad", "<i4", 2)])  # 80 B
TEXDESC_DTYPE = np.dtype([("offset", "<u4"), ("width", "<u4"), ("height", "<u4")])  # 12 B
assert NODE_DTYPE.itemsize == 48 and TRIANGLE_DTYPE.itemsize == 160 and MATERIAL_DTYPE.itemsize == 80 and TEXDESC_DTYPE.itemsize == 12


class SLOT:  # GPUTaskState SoA slot numbers (geom.h:199-236; FLX_S_* in fluctus_b200.h)
    ORIG, DIR, SHADOW_ORIG, SHADOW_DIR, T, EI, LAST_BSDF, LAST_EMISSION, LAST_T, P, N, UV = 0, 4, 8, 12, 16, 20, 24, 28, 32, 36, 40, 44
    PHASE, LAST_PDF_W, PATH_LEN, SEED, LAST_SPECULAR, SHADOW_BLOCKED, BACKFACE, PIXEL_INDEX, FIRST_DIFFUSE = 46, 47, 48, 49, 50, 51, 52, 53, 54
    LAST_PDF_DIRECT, LAST_PDF_IMPLICIT, LAST_COS_TH, LAST_LIGHT_PICK, SHADOW_RAY_LEN, HIT_T, HIT_I, AREA_LIGHT_HIT, MAT_ID = 55, 56, 57, 58, 59, 60, 61, 62, 63
    COUNT = 64
    # float3 padding lanes and the microkernel-only phase field are never written by the wavefront path
    UNUSED = (3, 7, 11, 15, 19, 23, 27, 31, 35, 39, 43, 46)
    NAMES = {0: "orig", 4: "dir", 8: "shadowOrig", 12: "shadowDir", 16: "T", 20: "Ei", 24: "lastBsdf", 28: "lastEmission", 32: "lastT", 36: "P",
             40: "N", 44: "uvTex", 47: "lastPdfW", 48: "pathLen", 49: "seed", 50: "lastSpecular", 51: "shadowRayBlocked", 52: "backfaceHit",
             53: "pixelIndex", 54: "firstDiffuseHit", 55: "lastPdfDirect", 56: "lastPdfImplicit", 57: "lastCosTh", 58: "lastLightPickProb",
             59: "shadowRayLen", 60: "t", 61: "i", 62: "areaLightHit", 63: "matId"}


class BXDF:  # bxdf_types.h:4-11
    DIFFUSE, GLOSSY, GGX_ROUGH_REFLECTION, IDEAL_REFLECTION, GGX_ROUGH_DIELECTRIC, IDEAL_DIELECTRIC, EMISSIVE = 2, 4, 8, 16, 32, 64, 128


QUEUE_NAMES = ("raygen", "extension", "shadow", "diffuse", "glossy", "ggxRefl", "ggxRefr", "delta")
