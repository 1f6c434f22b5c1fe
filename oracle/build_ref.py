#!/usr/bin/env python
"""build_ref.py -- TEST INFRASTRUCTURE (oracle).

Builds oracle/_ref/ from the reference tree where it lies (default /root/reference, override
with FLX_REFERENCE_DIR):

  oracle/_ref/libfluctus_ref.so   the reference's wavefront kernels (src/wf_*.cl and what they
                                  include) compiled for the host through oracle/ref_shim/cl_shim.hpp,
                                  serial (ref_*) and OpenMP (ref_par_*) entry points
  oracle/_ref/scene_tool          the reference's OBJ/PLY import + SBVH builder + env-map tables

Only binaries land in oracle/_ref/ (git-ignored, not gpurun-ignored).  The one textual transform
the kernel sources need -- OpenCL vector literals "(float3)(" -> C++ "float3(" -- is applied to
a copy in a temporary directory that is deleted afterwards; no reference source enters the repo.
The reference's own build system (cmake + OpenCL + GL + DevIL) is not used: none of those
dependencies exist in this image.
"""
import os
import re
import shutil
import subprocess
import sys
import tempfile
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SHIM = os.path.join(HERE, "ref_shim")
OUT = os.path.join(HERE, "_ref")
INC = os.path.join(os.path.dirname(HERE), "include")

KERNEL_TUS = {
    # name -> (REF_TU define, extra -D flags = the reference's per-kernel build options,
    #          src/kernel_impl.hpp:49-67 and 261-266 / src/utils.cpp:93-113, all switched on:
    #          every one of them is also guarded by the matching RenderParams field at run time)
    "reset": ("REF_TU_RESET", []),
    "raygen": ("REF_TU_RAYGEN", []),
    "ext": ("REF_TU_EXT", []),
    "shadow": ("REF_TU_SHADOW", []),
    "logic_single": ("REF_TU_LOGIC_SINGLE", ["USE_AREA_LIGHT", "USE_ENV_MAP", "SAMPLE_EXPLICIT", "SAMPLE_IMPLICIT", "WF_SINGLE_MAT_QUEUE"]),
    "logic_separate": ("REF_TU_LOGIC_SEPARATE", ["USE_AREA_LIGHT", "USE_ENV_MAP", "SAMPLE_EXPLICIT", "SAMPLE_IMPLICIT"]),
    "mat_all": ("REF_TU_MAT_ALL", ["BXDF_USE_DIFFUSE", "BXDF_USE_GLOSSY", "BXDF_USE_GGX_ROUGH_REFLECTION", "BXDF_USE_IDEAL_REFLECTION",
                                   "BXDF_USE_GGX_ROUGH_DIELECTRIC", "BXDF_USE_IDEAL_DIELECTRIC", "BXDF_USE_EMISSIVE"]),
    "mat_diffuse": ("REF_TU_MAT_DIFFUSE", []),
    "mat_glossy": ("REF_TU_MAT_GLOSSY", []),
    "mat_ggx_refl": ("REF_TU_MAT_GGX_REFL", []),
    "mat_ggx_refr": ("REF_TU_MAT_GGX_REFR", []),
    "mat_delta": ("REF_TU_MAT_DELTA", []),
    "postprocess": ("REF_TU_POSTPROCESS", []),  # src/mk_postprocess.cl + src/tonemap.cl (the display pass of the render loop, tracer.cpp:302,447)
    # the microkernel integrator (src/mk_*.cl; clcontext.cpp:709-750).  sampleBsdf gets every BXDF_USE_* the scene could
    # need (src/kernel_impl.hpp:376-387; each is also selected by material->type at run time)
    "mk_reset": ("REF_TU_MK_RESET", []),
    "mk_raygen": ("REF_TU_MK_RAYGEN", []),
    "mk_next_vertex": ("REF_TU_MK_NEXT_VERTEX", []),
    "mk_sample_bsdf": ("REF_TU_MK_SAMPLE_BSDF", ["BXDF_USE_DIFFUSE", "BXDF_USE_GLOSSY", "BXDF_USE_GGX_ROUGH_REFLECTION", "BXDF_USE_IDEAL_REFLECTION",
                                                 "BXDF_USE_GGX_ROUGH_DIELECTRIC", "BXDF_USE_IDEAL_DIELECTRIC", "BXDF_USE_EMISSIVE"]),
    "mk_splat": ("REF_TU_MK_SPLAT", []),
    "mk_splat_preview": ("REF_TU_MK_SPLAT_PREVIEW", []),
}

WF = ["reset", "raygen", "ext", "shadow", "logic_single", "logic_separate", "mat_all", "mat_diffuse", "mat_glossy", "mat_ggx_refl", "mat_ggx_refr", "mat_delta", "postprocess"]
# Variants of the build, for the self-consistency pins of SURVEY 8(c) and the denoiser feature buffers.  Each: symbol infix ->
# (kernels, defines added, defines removed, also the OpenMP flavour?).  The base build ("" infix) is every kernel with
# -DUSE_SOA, the explicit-stack traversal and include/flx_math.h.
VARIANTS = {
    # the reference's kernels as built with Tracer::useDenoiser (src/kernel_impl.hpp:53,346,380,443)
    "dn_": (["logic_single", "logic_separate", "mk_next_vertex", "mk_sample_bsdf", "postprocess"], ["USE_OPTIX_DENOISER"], [], False),
    # stackless "bit stack" traversal (src/bvh.cl:10-230; Settings::getUseBitstack, src/clcontext.cpp:147): must find the same hits
    "bs_": (["ext", "shadow"], ["USE_BITSTACK"], [], False),
    # array-of-structures path state (no -DUSE_SOA: the macros of src/geom.h:26-37): must compute the same state
    "aos_": (WF, [], ["USE_SOA"], False),
    # the C library's sinf / cosf / ... instead of include/flx_math.h (cl_shim.hpp SHIM_LIBM): the independent-math oracle
    "lm_": (WF, ["SHIM_LIBM"], [], True),
}

VEC_LITERAL = re.compile(r"\((v?float[234]|int2)\)\(")


def reference_dir():
    return os.environ.get("FLX_REFERENCE_DIR", "/root/reference")


def available():
    return os.path.isfile(os.path.join(reference_dir(), "src", "wf_logic.cl"))


def _transform(src_dir, dst_dir):
    for name in os.listdir(src_dir):
        if not (name.endswith(".cl") or name in ("geom.h", "bxdf_types.h")):
            continue
        text = open(os.path.join(src_dir, name), encoding="utf-8", errors="replace").read()
        text = VEC_LITERAL.sub(r"\1(", text)
        if name == "ggx.cl":
            # The two rand() calls inside one vector literal are unsequenced in C++; OpenCL C
            # compilers (clang front end) evaluate them left to right.  Brace-init pins that order.
            text = text.replace("float2(rand(seed), rand(seed))", "float2{rand(seed), rand(seed)}")
        if name == "utils.cl":
            # OpenCL C converts float* -> volatile float4* implicitly (utils.cl:299,308); C++ needs the cast spelled out
            text = text.replace("return atomic_add_float3(ptr, value);", "return atomic_add_float3((volatile float3*)ptr, value);")
            text = text.replace("return atomic_add_float4(ptr, value);", "return atomic_add_float4((volatile float4*)ptr, value);")
        open(os.path.join(dst_dir, name), "w").write(text)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build_ref: command failed")


def build(force=False):
    if not available():
        raise RuntimeError("reference tree not found at %s" % reference_dir())
    lib = os.path.join(OUT, "libfluctus_ref.so")
    tool = os.path.join(OUT, "scene_tool")
    srcs = [os.path.join(SHIM, f) for f in ("cl_shim.hpp", "ref_kernels.cpp", "ref_abi.h", "scene_tool.cpp")] + [os.path.join(INC, "flx_math.h"), __file__]
    newest = max(os.path.getmtime(s) for s in srcs)
    if not force and all(os.path.exists(p) and os.path.getmtime(p) >= newest for p in (lib, tool)):
        return lib, tool
    os.makedirs(OUT, exist_ok=True)
    ref = reference_dir()
    tmp = tempfile.mkdtemp(prefix="flx_ref_")
    try:
        gen = os.path.join(tmp, "gen")
        os.makedirs(gen)
        _transform(os.path.join(ref, "src"), gen)
        common = ["g++", "-std=gnu++17", "-O2", "-fPIC", "-fpermissive", "-ffp-contract=off", "-fno-fast-math", "-w",
                  "-DGPU", "-DUSE_SOA", "-DAPPLE_SILICON", "-I", SHIM, "-I", INC, "-I", gen]
        jobs = []
        objs = []

        def add(name, variant, extra, drop, par):
            tu, defs = KERNEL_TUS[name]
            obj = os.path.join(tmp, "%s%s%s.o" % (variant, name, "_par" if par else ""))
            cmd = [c for c in common if c not in ["-D" + d for d in drop]] + ["-D" + tu] + ["-D" + d for d in defs + extra]
            if variant:
                cmd.append("-DREF_VARIANT=" + variant)
            if par:
                # the reference builds with -DFLT_FLOAT_ATOMICS (clcontext.cpp:145); needed once work-items run concurrently
                cmd += ["-DSHIM_PARALLEL", "-DFLT_FLOAT_ATOMICS", "-fopenmp", "-O3", "-march=x86-64-v3"]
            cmd += ["-c", os.path.join(SHIM, "ref_kernels.cpp"), "-o", obj]
            jobs.append(cmd)
            objs.append(obj)

        for name in KERNEL_TUS:
            for par in (False, True):
                add(name, "", [], [], par)
        for variant, (kernels, extra, drop, with_par) in VARIANTS.items():
            for name in kernels:
                add(name, variant, extra, drop, False)
                if with_par:
                    add(name, variant, extra, drop, True)
        with ThreadPoolExecutor(max_workers=os.cpu_count() or 4) as ex:
            list(ex.map(_run, jobs))
        _run(["g++", "-shared", "-fopenmp", "-o", lib] + objs)
        host = ["g++", "-std=c++14", "-O2", "-w", "-I", os.path.join(SHIM, "stubs"), "-I", os.path.join(ref, "include"), "-I", os.path.join(ref, "src"),
                os.path.join(SHIM, "scene_tool.cpp")] + [os.path.join(ref, "src", f) for f in ("bvh.cpp", "sbvh.cpp", "bvhnode.cpp", "envmap.cpp", "rgbe/rgbe.cpp")]
        _run(host + ["-o", tool])
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    return lib, tool


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
