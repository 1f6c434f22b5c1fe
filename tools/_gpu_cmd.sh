#!/bin/bash
set -x
mkdir -p gpurun_out
FLX_DEBUG_TIMING=1 timeout 300 python tools/e2e_breakdown.py 2>&1 | tail -24
