#!/bin/bash
set -x
mkdir -p gpurun_out
C="--variants 1 --thresholds 16 --ext-blocks 9 --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0"
timeout 900 python tools/tune_trace.py $C --overlaps 0,1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('overlap',r['overlap'],'ms/iter %.4f Mrays %.1f ext %.4f shadow %.4f'%(r['ms_per_iter'],r['mrays'],r['ext_ms'],r['shadow_ms']))"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_now.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_now.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -3
