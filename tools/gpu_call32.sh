#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call32.log 2>&1
echo "=== bench c5"
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench32_c5.json 2> gpurun_out/bench32_c5.err; echo "bench exit=$?"; tail -3 gpurun_out/bench32_c5.err; cat gpurun_out/bench32_c5.json
echo "=== done"
