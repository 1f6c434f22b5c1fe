#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "multi or gather or two" 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/bench2.err | grep '^{' | tail -1 > gpurun_out/bench_r2_2gpu_now.json; python -c "
import json; d=json.load(open('gpurun_out/bench_r2_2gpu_now.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['e2e']['value'], {k:(v['value'],v['gather_ms_per_call']) for k,v in d['gather'].items() if isinstance(v,dict)})"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 2 --warmup 1 --impl reference 2>/dev/null | grep '^{' | tail -1 | cut -c1-300
