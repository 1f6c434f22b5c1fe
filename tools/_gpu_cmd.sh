#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/run_configs.py --configs metric,C3,C4,C2 --iters 100 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); k=r['kernel_ms_per_iteration']; print('merged', r['config'], 'Mrays %.1f ms/iter %.4f ext %.4f shadow %.4f logic %.4f mat %.4f'%(r['mrays_per_s'], r['ms_per_iteration'], k['extrays'], k['shadowrays'], k['logic_fused'], k.get('materials',0)))"
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py tests/test_gpu_edge_cases.py -x -q 2>&1 | tail -2
