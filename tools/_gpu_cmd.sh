#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for t in "logic_tile=256" "logic_tile=128" "shadow_left_first=1" "logic_tile=128,shadow_left_first=1" "logic_tile=256"; do
  echo "== $t"
  timeout 300 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline --min-seconds 0.4 --tune $t | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],4), 'ext', r['avg_launch_ms'], 'shadow', r['shadow']['avg_launch_ms'], r['kernel_share_of_step'])"
done
for t in "overlap_trace=0" "overlap_trace=0,shadow_left_first=1"; do
  echo "== $t"
  timeout 300 python bench.py --steps 40 --warmup 10 --no-e2e --no-cpu-baseline --min-seconds 0.3 --tune $t | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']
print(round(d['value'],1), round(d['ms_per_step'],4), 'ext', r['avg_launch_ms'], 'shadow', r['shadow']['avg_launch_ms'], r['kernel_share_of_step'])"
done
