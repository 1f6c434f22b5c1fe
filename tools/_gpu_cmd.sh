#!/bin/bash
# GPU job of the moment (edited per gpurun call; kept for the record of what was run)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py tests/test_gpu_mk.py -x -q 2>&1 | tail -15
for t in "trace_variant=3" "trace_variant=1" "trace_variant=1,inner_bias=4" "trace_variant=1,inner_bias=8" "trace_variant=1,inner_bias=-4" "trace_variant=1,fetch_threshold=20" "trace_variant=1,fetch_threshold=24" "trace_variant=1,fetch_threshold=12" "trace_variant=1,ext_min_blocks=8" "trace_variant=1,ext_min_blocks=10" "trace_variant=1,overlap_trace=0" "trace_variant=3,overlap_trace=0"; do
  echo "== $t"
  timeout 300 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline --min-seconds 0.2 --tune $t | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],4), 'ext', r['avg_launch_ms'], 'shadow', r['shadow']['avg_launch_ms'], r['kernel_share_of_step'])"
done
timeout 600 python -m pytest tests/test_gpu_parity_large.py -x -q 2>&1 | tail -5
