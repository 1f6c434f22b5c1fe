#!/bin/bash
set -x
mkdir -p gpurun_out
C="--variants 1 --thresholds 16 --ext-blocks 9 --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0"
timeout 900 python tools/tune_trace.py $C --overlaps 0 --hot-nodes 0,512,1024,2048,4096,8192,1,0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('hot',r['hot_nodes'],'ms/iter %.4f Mrays %.1f ext %.4f shadow %.4f'%(r['ms_per_iter'],r['mrays'],r['ext_ms'],r['shadow_ms']))"
