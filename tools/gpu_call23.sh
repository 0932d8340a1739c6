#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call23.log 2>&1
echo "=== pytest backward"
timeout 900 python -m pytest tests/test_backward_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -3
echo "=== train profile c4"
timeout 300 python tools/train_profile.py --workload c4
echo "=== done"
