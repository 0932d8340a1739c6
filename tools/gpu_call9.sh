#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call9.log 2>&1
P="python tools/tc_probe.py"
echo "=== correctness"
timeout 120 $P --m 4096 --n 1152 --k 1152 --cfgs 2x256,2x192,2x128,1x256,1x128 --check --iters 5 --no-cublas
timeout 120 $P --m 4000 --n 1000 --k 1152 --cfgs 2x256,1x192 --check --iters 5 --no-cublas --epi 2
timeout 120 $P --m 4096 --n 1000 --k 1152 --cfgs 2x256 --check --iters 5 --no-cublas --epi 1
echo "=== epilogue-only cost (K=64)"
timeout 120 $P --m 16384 --n 1152 --k 64 --cfgs 2x256,2x128 --epi 2 --no-cublas
timeout 120 $P --m 16384 --n 3456 --k 64 --cfgs 2x256,2x128 --no-cublas
timeout 120 $P --m 16384 --n 4608 --k 64 --cfgs 2x256,2x128 --epi 1 --no-cublas
echo "=== C3 shapes"
timeout 120 $P --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x192,2x128 --epi 2
timeout 120 $P --m 16384 --n 1152 --k 4608 --cfgs 2x256,2x192,2x128 --epi 2
timeout 120 $P --m 16384 --n 3456 --k 1152 --cfgs 2x256,2x192,2x128
timeout 120 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,2x192,2x128 --epi 1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -x 2>&1 | grep -v "^$" | tail -5
echo "=== done"
