#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/_tree_try.py conference sp0.01_ri16,sp0.003_ri16,sp0.001_ri16,sp0.0003_ri16,sp0.0001_ri16 2>&1 | tail -3
