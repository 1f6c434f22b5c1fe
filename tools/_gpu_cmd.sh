#!/bin/bash
mkdir -p gpurun_out
for n in default any4 any5; do
  if [ $n = default ]; then unset FLX_LIB_PATH; else export FLX_LIB_PATH=$PWD/fluctus_b200/_ab/lib_$n.so; fi
  for sc in conference country_kitchen luxball; do
  timeout 300 python tools/tune_trace.py --scene $sc --thresholds 16 --ext-blocks 9 --shadow-blocks 10 --max-l1 0 --iters 40 --smem-stacks 0 --variants 1 --inner-mins 8 --overlaps 0 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('$n $sc', 'ms/iter %.4f ext %.4f shadow %.4f'%(r['ms_per_iter'], r['ext_ms'], r['shadow_ms']))"
  done
done
