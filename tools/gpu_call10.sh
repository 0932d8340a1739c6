#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call10.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== failing test"
timeout 600 python -m pytest tests/test_backward_gpu.py -q -m gpu -k "data_parallel_wrapper" -p no:cacheprovider -x 2>&1 | tail -30
echo "=== epilogue-only (K=64) flags 0/1/2/3"
for d in 0 1 2 3; do
echo "--- debug=$d"
DITB200_EPI_DEBUG=$d timeout 120 $P --m 16384 --n 3456 --k 64 --cfgs 2x256
DITB200_EPI_DEBUG=$d timeout 120 $P --m 16384 --n 4608 --k 64 --cfgs 2x256 --epi 1
DITB200_EPI_DEBUG=$d timeout 120 $P --m 16384 --n 1152 --k 64 --cfgs 2x256 --epi 2
done
echo "=== fc1 with tanh.approx"
timeout 120 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,2x192 --epi 1 --check
echo "=== done"
