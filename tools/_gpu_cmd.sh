#!/bin/bash
mkdir -p gpurun_out
p=29550
for ch in default 2 4 16; do
  p=$((p+1))
  if [ $ch = default ]; then unset NCCL_MAX_NCHANNELS; unset NCCL_MIN_NCHANNELS; else export NCCL_MAX_NCHANNELS=$ch; fi
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $p bench.py --gpus 2 --steps 20 --warmup 5 --config c5 --no-cpu-baseline --no-e2e 2>/dev/null | grep '^{' | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); g=d['gather']; print('channels $ch', d['value'], {k:(v['value'],v['gather_ms_per_call']) for k,v in g.items() if isinstance(v,dict)})"
done
