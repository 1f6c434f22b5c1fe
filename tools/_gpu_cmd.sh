#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python tools/run_configs.py --configs metric,C2 --iters 200 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('fold', r['config'], 'Mrays %.1f ms/iter %.4f'%(r['mrays_per_s'], r['ms_per_iteration']))"
done
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_edge_cases.py tests/test_gpu_mk.py -x -q 2>&1 | tail -2
