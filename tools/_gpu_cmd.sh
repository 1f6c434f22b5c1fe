#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_now.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_now.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['runs'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_persistent|k_logic' -s 36 -c 3 -f -o gpurun_out/r2_final_prof python tools/prof_step.py 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python tools/prof_step.py --warmup 12 --iters 2 2>&1 | tail -1
timeout 300 python tools/run_configs.py > gpurun_out/r2_configs_single_gpu.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2_configs_single_gpu.jsonl'):
    r=json.loads(l); print(r['config'], r['mrays_per_s'], r['ms_per_iteration'])"
timeout 300 python tools/bench_mk.py > gpurun_out/r2_mk_integrator.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2_mk_integrator.jsonl'):
    r=json.loads(l); print(r.get('integrator'), r['scene'], r['mrays_per_s'])"
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
