#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_r2_now.json 2> gpurun_out/bench_r2_now.err; tail -c 2500 gpurun_out/bench_r2_now.json
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
