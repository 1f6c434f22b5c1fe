#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_now.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_now.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['e2e']['runs'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_step.py --warmup 12 --iters 2 2>&1 | tail -1
