#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call18.log 2>&1
echo "=== pytest kernels"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_parity_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench18_c3.json 2> gpurun_out/bench18_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench18_c3.err; cat gpurun_out/bench18_c3.json
echo "=== done"
