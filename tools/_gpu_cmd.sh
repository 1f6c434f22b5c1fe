#!/bin/bash
# GPU job of the moment (edited per gpurun call; kept for the record of what was run)
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
nproc
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err; tail -c 600 gpurun_out/r2_bench_a.err
python - <<'P'
import json
d=json.load(open('gpurun_out/r2_bench_a.json'))
print({k:d[k] for k in ('value','ms_per_step','blocks','e2e','clocks')})
print(d['roofline']['kernel_share_of_step'], d['roofline']['avg_launch_ms'], d['roofline']['shadow']['avg_launch_ms'], d['cpu_baseline'])
P
FLX_DEBUG_TIMING=1 timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -12
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_base_launches.csv python tools/prof_step.py > gpurun_out/r2_base_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_persistent|k_logic' -s 24 -c 3 -o gpurun_out/r2_base_prof python tools/prof_step.py > gpurun_out/r2_base_prof.log 2>&1
tail -3 gpurun_out/r2_base_prof.log
