#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity_large.py -x -q -k "luxball or kitchen" 2>&1 | tail -3
for t in "material_mask=1" "material_mask=0"; do
timeout 300 python tools/run_configs.py --configs C3,C4 --iters 100 --tune $t 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); k=r['kernel_ms_per_iteration']; print('$t', r['config'], 'Mrays %.1f ms/iter %.4f ext %.4f shadow %.4f logic %.4f mat %.4f'%(r['mrays_per_s'], r['ms_per_iteration'], k['extrays'], k['shadowrays'], k['logic_fused'], k.get('materials',0)))"
done
