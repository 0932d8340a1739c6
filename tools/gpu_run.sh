#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run.log 2>&1
echo "=== pytest backward + kernels"
timeout 1200 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== train profile c4"
timeout 300 python tools/train_profile.py --workload c4 > gpurun_out/tp.log 2>&1; head -22 gpurun_out/tp.log
