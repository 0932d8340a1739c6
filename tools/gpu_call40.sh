#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call40.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== correctness"
timeout 60 $P --m 4096 --n 1152 --k 1152 --cfgs 0x0,2x256,2x192,2x128,1x256 --check --iters 3 --epi 2 --inplace; echo "exit=$?"
timeout 60 $P --m 3840 --n 1000 --k 1152 --cfgs 0x0,1x192 --check --iters 3 --epi 2; echo "exit=$?"
echo "=== timing"
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 0x0,2x256 --epi 2 --inplace
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 0x0,2x256 --epi 2 --inplace
echo "=== done"
