#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bvh_build.py -x -q 2>&1 | tail -15
