#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call30.log 2>&1
P="python tools/tc_probe.py --no-cublas"
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 0x0 --check
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1 --check
timeout 90 $P --m 16384 --n 3456 --k 64 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 64 --cfgs 0x0 --epi 1
timeout 90 $P --m 4000 --n 1000 --k 1152 --cfgs 0x0,1x128 --epi 1 --check
echo "=== pytest backward + kernels"
timeout 1200 python -m pytest tests/test_backward_gpu.py tests/test_kernels_gpu.py -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== train profile c4"
timeout 300 python tools/train_profile.py --workload c4 > gpurun_out/tp30.log 2>&1; head -22 gpurun_out/tp30.log
echo "=== done"
