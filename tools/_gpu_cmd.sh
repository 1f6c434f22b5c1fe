#!/bin/bash
# GPU job of the moment (edited per gpurun call; kept for the record of what was run)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_aov_and_pins.py -x -q 2>&1 | tail -15
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_mk.py -x -q 2>&1 | tail -5
for t in "gather_priority=0" "gather_priority=1"; do
  echo "== c5 2 gpus $t"
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus 2 --config c5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --min-seconds 0.3 --tune $t 2>/dev/null | grep '^{' | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print(round(d['value'],1), d['gather'])"
done
