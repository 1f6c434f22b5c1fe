#!/bin/bash
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2_san_memcheck.log 2>&1; tail -2 gpurun_out/r2_san_memcheck.log
timeout 1500 compute-sanitizer --tool initcheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2_san_initcheck.log 2>&1; tail -2 gpurun_out/r2_san_initcheck.log
timeout 1500 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize.py > gpurun_out/r2_san_racecheck.log 2>&1; tail -2 gpurun_out/r2_san_racecheck.log
