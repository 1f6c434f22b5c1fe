#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/run_configs.py > gpurun_out/r2_configs_single_gpu.jsonl 2>/dev/null; python -c "
import json
for l in open('gpurun_out/r2_configs_single_gpu.jsonl'):
    r=json.loads(l); print(r['config'], r['mrays_per_s'], r['ms_per_iteration'])"
