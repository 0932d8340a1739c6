#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call50.log 2>&1
P="python tools/tc_probe.py --no-cublas --trans-a --trans-w --iters 20"
for shape in "1152 4608" "4608 1152" "3456 1152" "1152 1152"; do
set -- $shape
for sk in 1 2 3 4 6; do
timeout 90 $P --m $1 --n $2 --k 8192 --cfgs 2x256,2x128,1x256,1x128 --split-k $sk 2>&1 | grep -v "^$"
done
done
