#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge_cases.py -x -q -k "multi_gpu or gather" 2>&1 | tail -6
for t in "gather_direct=0" "gather_direct=1" "gather_direct=1,gather_priority=1"; do
  echo "== c5 2 gpus $t"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --min-seconds 0.3 --tune $t 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); g=d['gather']; print(round(d['value'],1), {k:(v['value'],v['ms_per_step'],v['gather_ms_per_call']) for k,v in g.items() if isinstance(v,dict)})"
done
