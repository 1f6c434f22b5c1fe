#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/bench8.err | grep '^{' | tail -1 > gpurun_out/bench_r2_8gpu_now.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --steps 20 --warmup 5 --config c5 2>gpurun_out/bench8c5.err | grep '^{' | tail -1 > gpurun_out/bench_r2_c5_8gpu_now.json
python -c "
import json
for f in ('gpurun_out/bench_r2_8gpu_now.json','gpurun_out/bench_r2_c5_8gpu_now.json'):
    d=json.load(open(f)); print(f, {k:d[k] for k in ('value','n_gpus','ms_per_step')}, 'e2e', d['e2e']['value'], {k:(v['value'],v['gather_ms_per_call']) for k,v in d['gather'].items() if isinstance(v,dict)})"
tail -2 gpurun_out/bench8.err
