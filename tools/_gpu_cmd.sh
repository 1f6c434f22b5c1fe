#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/run_benchmark.py --seconds 3 --integrator wf --out gpurun_out/r2_benchmark_protocol_wf.csv 2>&1 | tail -5 > gpurun_out/r2_benchmark_protocol_summary.txt
timeout 600 python tools/run_benchmark.py --seconds 3 --integrator mk --out gpurun_out/r2_benchmark_protocol_mk.csv 2>&1 | tail -5 >> gpurun_out/r2_benchmark_protocol_summary.txt
cat gpurun_out/r2_benchmark_protocol_summary.txt
