#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do
timeout 300 python tools/run_configs.py --configs metric,C2 --iters 200 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('fold2', r['config'], 'Mrays %.1f ms/iter %.4f'%(r['mrays_per_s'], r['ms_per_iteration']))"
done
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
