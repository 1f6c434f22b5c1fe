#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
FLX_DEBUG_TIMING=1 timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; grep -c upload gpurun_out/r2_bench_d.err
python -c "
import json
d=json.load(open('gpurun_out/r2_bench_d.json')); print({k:d[k] for k in ('value','ms_per_step','blocks')}); print(d['e2e']); print(d['roofline']['kernel_share_of_step'], d['cpu_baseline'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('reference arm', d['value'], d['cpu_baseline']['cores'])"
timeout 300 ./examples/flx_render_file oracle/_ref/assets/country_kitchen/Country-Kitchen.obj gpurun_out/r2_kitchen_from_files.png 640 360 32 6 oracle/_ref/assets/env_maps/night.hdr 30 2>&1 | tail -4
