#!/usr/bin/env python
"""build_oracle.py -- TEST INFRASTRUCTURE. Compiles the C restatement oracle/wf_oracle.c into oracle/liboracle.so
(serial port_* entry points = deterministic oracle; OpenMP port_par_* = CPU baseline when oracle/_ref is absent)."""
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "liboracle.so")
SRC = os.path.join(HERE, "wf_oracle.c")
BVH_SRC = os.path.join(HERE, "bvh_oracle.c")  # CPU restatement of the GPU hierarchy builder
INC = os.path.join(os.path.dirname(HERE), "include")


def build(force=False):
    deps = [SRC, BVH_SRC, os.path.join(INC, "flx_math.h"), os.path.join(HERE, "ref_shim", "ref_abi.h"), __file__]
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= max(os.path.getmtime(d) for d in deps):
        return OUT
    with tempfile.TemporaryDirectory() as tmp:
        common = ["gcc", "-std=gnu11", "-fPIC", "-ffp-contract=off", "-fno-fast-math", "-Wall", "-Wno-unused-function", "-Wno-unused-variable",
                  "-I", INC, "-I", HERE, "-c", SRC]
        objs = [os.path.join(tmp, "serial.o"), os.path.join(tmp, "par.o"), os.path.join(tmp, "bvh.o")]
        cmds = [common + ["-O2", "-o", objs[0]], common + ["-O3", "-march=x86-64-v3", "-fopenmp", "-DPORT_PARALLEL", "-o", objs[1]],
                common[:-1] + [BVH_SRC, "-O2", "-o", objs[2]]]
        for c in cmds + [["gcc", "-shared", "-fopenmp", "-o", OUT] + objs + ["-lm"]]:
            r = subprocess.run(c, capture_output=True, text=True)
            if r.returncode != 0:
                sys.stderr.write(" ".join(c) + "\n" + r.stdout + r.stderr)
                raise RuntimeError("build_oracle: gcc failed")
            if r.stderr.strip():
                sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
