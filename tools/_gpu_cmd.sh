#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_persistent|k_logic' -s 36 -c 3 -f -o gpurun_out/r2_final_prof python tools/prof_step.py 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_step.py --warmup 12 --iters 2 2>&1 | tail -2
