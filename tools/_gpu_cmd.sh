#!/bin/bash
set -x
mkdir -p gpurun_out
C="--thresholds 16 --ext-blocks 9 --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0"
timeout 900 python tools/tune_trace.py $C --variants 1,3 --overlaps 0,1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('variant',r['variant'],'overlap',r['overlap'],'ms/iter %.4f Mrays %.1f ext %.4f shadow %.4f'%(r['ms_per_iter'],r['mrays'],r['ext_ms'],r['shadow_ms']))"
for eb in 8 9 10; do timeout 300 python tools/tune_trace.py --thresholds 12,16,24 --ext-blocks $eb --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0 --variants 3 --overlaps 1 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('variant',r['variant'],'thr',r['threshold'],'ext_blocks',r['ext_blocks'],'ms/iter %.4f Mrays %.1f ext %.4f shadow %.4f'%(r['ms_per_iter'],r['mrays'],r['ext_ms'],r['shadow_ms']))"; done
