#!/bin/bash
# Scratch job script for `gpurun -- 'bash tools/_gpu_cmd.sh'` (one GPU call = one edit of this file).  The standard end-of-change check:
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_now.json
timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
