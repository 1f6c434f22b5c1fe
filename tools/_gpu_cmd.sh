#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_bvh_build.py > gpurun_out/r2_gpu_bvh_build_opt.jsonl 2> gpurun_out/bvh_build.err; tail -3 gpurun_out/bvh_build.err
python - <<'PY'
import json
for l in open('gpurun_out/r2_gpu_bvh_build_opt.jsonl'):
    r=json.loads(l)
    print(r['scene'], {k:v for k,v in r.items() if k.startswith('build_ms') or k.startswith('throughput')})
    for k,v in r.items():
        if isinstance(v,dict): print('   ',k,v)
PY
for it in 4 8 16 32; do python - <<PY
import sys,os
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from fluctus_b200 import CLContext, SceneData
ref=SceneData.load_blob('oracle/_ref/scenes/conference.bin')
with CLContext(1024) as c:
    c.setTuning(bvh_reinsert=$it)
    c.buildBVH(ref.tris,8,'ploc_opt')
    print('iterations',$it,'build ms',min(c.buildBVH(ref.tris,8,'ploc_opt')[2] for _ in range(3)))
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'k_ri_|k_ploc' -c 400 --csv --log-file gpurun_out/r2_bvh_opt_launches.csv python -c "
import sys
sys.path.insert(0,'.'); sys.path.insert(0,'tests')
from fluctus_b200 import CLContext, SceneData
ref=SceneData.load_blob('oracle/_ref/scenes/conference.bin')
with CLContext(1024) as c:
    c.buildBVH(ref.tris,8,'ploc_opt')
" > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(l for l in open('gpurun_out/r2_bvh_opt_launches.csv') if l.startswith('"')))
h=rows[0]; agg=collections.defaultdict(lambda:[0,0.0])
for r in rows[1:]:
    n=r[h.index('Kernel Name')].split('(')[0]; v=float(r[h.index('Metric Value')]); u=r[h.index('Metric Unit')]
    v = v/1e3 if u in ('ns','nsecond') else v
    agg[n][0]+=1; agg[n][1]+=v
for n,(c,t) in sorted(agg.items(), key=lambda kv:-kv[1][1]): print('%-40s launches %4d total %10.1f us'%(n,c,t))
PY
