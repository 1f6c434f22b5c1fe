#!/bin/bash
set -x
mkdir -p gpurun_out
C="--thresholds 16 --ext-blocks 9 --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0 --variants 1"
timeout 900 python tools/tune_trace.py $C --overlaps 0,1 --inner-mins 8 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('noknob inner_min',r['inner_min'],'overlap',r['overlap'],'ms/iter %.4f Mrays %.1f ext %.4f shadow %.4f'%(r['ms_per_iter'],r['mrays'],r['ext_ms'],r['shadow_ms']))"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mk.py -x -q 2>&1 | tail -2
