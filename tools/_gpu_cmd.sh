#!/bin/bash
set -x
mkdir -p gpurun_out
C="--thresholds 12,16,20,24 --ext-blocks 8,9,10 --shadow-blocks 9,10 --max-l1 0 --iters 40 --smem-stacks 0 --variants 1 --inner-mins 6,8,10"
timeout 900 python tools/tune_trace.py $C --overlaps 1 > gpurun_out/resweep2.jsonl 2>/dev/null
python - <<'PY'
import json
rows=[json.loads(l) for l in open('gpurun_out/resweep2.jsonl')]
rows.sort(key=lambda r:r['ms_per_iter'])
for r in rows[:10]:
    print('thr %2d inner_min %2d ext_blocks %2d shadow_blocks %2d  ms/iter %.4f Mrays %.1f'%(r['threshold'],r['inner_min'],r['ext_blocks'],r['shadow_blocks'],r['ms_per_iter'],r['mrays']))
base=[r for r in rows if r['threshold']==16 and r['inner_min']==8 and r['ext_blocks']==9 and r['shadow_blocks']==10]
print('current default:',base[0]['ms_per_iter'],base[0]['mrays'])
PY
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 2>/dev/null | grep '^{' | tail -1 > gpurun_out/bench_r2_now.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_now.json')); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['e2e']['runs'], d['roofline']['avg_launch_ms'], d['roofline']['frac'])"
