#!/usr/bin/env python
"""SASS evidence for profiles/: per kernel of libfluctus_b200.so, the static instruction count and the instructions that show what it is
built from (256-bit loads and their L1 eviction hints, bulk copies + mbarriers, global atomics / reductions, warp votes, local-memory
traffic), with the first occurrences as they appear in `cuobjdump -sass`.  Runs here (no GPU needed):

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.txt
"""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PATTERNS = [
    ("LDG.E.*256 (ld.global.nc.v8.f32, new on sm_100)", re.compile(r"\bLDG\.E(\.\w+)*\.256")),
    ("  of which .EL (L1 evict-last: inner nodes)", re.compile(r"\bLDG\.E\.EL(\.\w+)*\.256")),
    ("  of which .NA (no L1 allocation: leaf triangles, hit attributes)", re.compile(r"\bLDG\.E\.NA(\.\w+)*\.256")),
    ("LDG.E.*128", re.compile(r"\bLDG\.E(\.\w+)*\.128")),
    ("UBLKCP (cp.async.bulk)", re.compile(r"\bUBLKCP")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS\.")),
    ("ATOMG / RED (global atomics)", re.compile(r"\b(ATOMG|RED)\.")),
    ("ATOMS (shared atomics)", re.compile(r"\bATOMS\.")),
    ("VOTE (ballot)", re.compile(r"\bVOTEU?\.")),
    ("LDL / STL (local memory: traversal stack, spills)", re.compile(r"\b(LDL|STL)\b")),
    ("CCTL / prefetch", re.compile(r"\bCCTL")),
    ("tensor-core instructions (UTC*MMA / HMMA / IMMA)", re.compile(r"\b(UTC\w*MMA|HMMA|IMMA|QMMA)\b")),
]


def main():
    lib = os.path.join(ROOT, "fluctus_b200", "libfluctus_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    funcs, cur = [], None
    for line in sass.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = [names[len(funcs)], []]
            funcs.append(cur)
        elif cur is not None and re.search(r"/\*[0-9a-f]{4,}\*/\s+\S", line) and ";" in line:
            cur[1].append(line.rstrip())
    print("SASS evidence, cuobjdump -sass fluctus_b200/libfluctus_b200.so (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), tools/sass_excerpt.py.")
    print("Per kernel: static instruction count and how many of them are the instructions that show what the kernel is built from; below each, the")
    print("first occurrences as they appear in the listing.  No tensor-core instructions anywhere: the path is branchy fp32 traversal and SoA")
    print("streaming (BASELINE.json north_star).\n")
    total_tc = 0
    for name, ins in sorted(funcs, key=lambda f: f[0]):
        short = re.sub(r"\s+", " ", name)
        print("== %s\n   %d instructions" % (short[:230], len(ins)))
        for label, rx in PATTERNS:
            hits = [l for l in ins if rx.search(l)]
            if label.startswith("tensor-core"):
                total_tc += len(hits)
            if not hits:
                continue
            print("      %d x %s" % (len(hits), label))
            if not label.startswith("  of which"):
                for l in hits[:2]:
                    print("            " + re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l.strip()))
        print()
    print("tensor-core instructions in the whole library: %d" % total_tc)


if __name__ == "__main__":
    main()
