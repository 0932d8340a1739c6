#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call12.log 2>&1
P="python tools/tc_probe.py --no-cublas"
for nn in 0 1; do
if [ $nn = 1 ]; then export DITB200_NO_NARROW=1; fi
echo "=== NO_NARROW=$nn"
timeout 120 $P --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x192 --epi 2 --check
timeout 120 $P --m 16384 --n 1152 --k 4608 --cfgs 2x256,2x192 --epi 2
timeout 120 $P --m 16384 --n 3456 --k 1152 --cfgs 2x256,2x192
timeout 120 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,2x192 --epi 1
timeout 120 $P --m 8192 --n 1152 --k 1152 --cfgs 2x256,2x192 --epi 2
timeout 120 $P --m 8192 --n 1152 --k 4608 --cfgs 2x256,2x192 --epi 2
timeout 120 $P --m 8192 --n 3456 --k 1152 --cfgs 2x256,2x192
done
echo "=== done"
