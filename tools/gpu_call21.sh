#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call21.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== stream-K correctness"
timeout 60 $P --m 4096 --n 1152 --k 1152 --cfgs 0x0,2x256,2x192,2x128 --check --iters 3 --epi 2 --inplace; echo "exit=$?"
timeout 60 $P --m 1152 --n 1152 --k 4096 --cfgs 0x0,2x256,1x128 --check --iters 3 --trans-w --trans-a --split-k 4; echo "exit=$?"
timeout 60 $P --m 1000 --n 712 --k 8192 --cfgs 0x0 --check --iters 3 --split-k 4; echo "exit=$?"
for nn in 0 1; do
if [ $nn = 1 ]; then export DITB200_NO_STREAMK=1; fi
echo "=== NO_STREAMK=$nn"
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 0x0,2x256,2x192 --epi 2 --inplace
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 0x0,2x256,2x192 --epi 2 --inplace
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 0x0 --trans-w --trans-a --split-k 3
timeout 90 $P --m 4608 --n 1152 --k 8192 --cfgs 0x0 --trans-w --trans-a --split-k 3
timeout 90 $P --m 1152 --n 1152 --k 8192 --cfgs 0x0 --trans-w --trans-a --split-k 6
timeout 90 $P --m 3456 --n 1152 --k 8192 --cfgs 0x0 --trans-w --trans-a --split-k 2
done
echo "=== done"
