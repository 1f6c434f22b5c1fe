"""Builds fluctus_b200/libfluctus_b200.so with nvcc for sm_100a (in-tree, so the .so travels to the GPU box)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
# FLX_LIB_OUT / FLX_NVCC_EXTRA: A/B builds of an experiment next to the product library (tools/; the package loads FLX_LIB_PATH when set)
OUT = os.environ.get("FLX_LIB_OUT") or os.path.join(os.path.dirname(HERE), "libfluctus_b200.so")
SOURCES = ["flx_api.cu", "flx_scene_io.cpp", "flx_jpeg.cpp"]
DEPS = ["flx_api.cu", "flx_scene_io.cpp", "flx_jpeg.cpp", "flx_trace_greedy.cuh", "flx_kernels.cuh", "flx_mk.cuh", "flx_bvh_build.cuh", "flx_bvh_repack.cuh", "flx_trace.cuh", "flx_trace_persistent.cuh", "flx_bsdf.cuh", "flx_device.cuh", os.path.join(ROOT, "include", "fluctus_b200.h"),
        os.path.join(ROOT, "include", "flx_math.h"), "build.py"]
# -fmad=false + IEEE div/sqrt + no ftz: arithmetic is bit-identical to the host oracle (DESIGN.md "Numerics")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-fmad=false", "-prec-div=true", "-prec-sqrt=true",
         "-ftz=false", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared"]


def build(force=False, verbose=False):
    deps = [d if os.path.isabs(d) else os.path.join(HERE, d) for d in DEPS]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc] + FLAGS + os.environ.get("FLX_NVCC_EXTRA", "").split() + (["-Xptxas", "-v"] if verbose else []) + ["-I", os.path.join(ROOT, "include"), "-I", HERE, "-o", OUT] + \
          [os.path.join(HERE, s) for s in SOURCES] + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose:
        sys.stderr.write(r.stderr)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("nvcc failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
