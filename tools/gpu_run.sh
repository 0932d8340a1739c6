#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_c3_v7.json 2>gpurun_out/bench_err.log; tail -c 1500 gpurun_out/bench_c3_v7.json
