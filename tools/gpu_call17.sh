#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call17.log 2>&1
P="python tools/tc_probe.py"
echo "=== multicast correctness"
timeout 60 $P --m 4096 --n 1024 --k 1152 --cfgs 4x256 --check --iters 3 --no-cublas; echo "exit=$?"
timeout 60 $P --m 4000 --n 1152 --k 1152 --cfgs 4x256,4x128 --check --iters 3 --no-cublas --epi 2; echo "exit=$?"
echo "=== C3 shapes"
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x192,4x256,4x128 --epi 2 --no-cublas
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 2x256,2x192,4x256,4x128 --epi 2 --no-cublas
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 2x256,4x256 --no-cublas
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,4x256 --epi 1 --no-cublas
timeout 90 $P --m 18944 --n 4608 --k 4608 --cfgs 2x256,4x256
echo "=== done"
