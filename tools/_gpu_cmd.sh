#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:'k_trace_persistent' -s 24 -c 2 -f -o gpurun_out/r2_plocopt_prof python tools/prof_step.py --builder ploc_opt 2>&1 | tail -3
