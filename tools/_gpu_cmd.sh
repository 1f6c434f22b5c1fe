#!/bin/bash
set -x
mkdir -p gpurun_out
for t in "material_mask=0,fused_min_blocks=3" "material_mask=1,fused_min_blocks=3" "material_mask=0,fused_min_blocks=4" "material_mask=1,fused_min_blocks=4" "material_mask=1,fused_min_blocks=2" "material_mask=1,fused_min_blocks=3"; do
timeout 300 python tools/run_configs.py --configs metric --tune $t 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l); print('$t', r['mrays_per_s'], r['ms_per_iteration'], r['kernel_ms_per_iteration'])
"
done
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_large.py -x -q 2>&1 | tail -3
