#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call7.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== epilogue-only cost (K=64): time/waves = E"
timeout 120 $P --m 16384 --n 1152 --k 64 --cfgs 2x256,2x128 --epi 2
timeout 120 $P --m 16384 --n 3456 --k 64 --cfgs 2x256,2x128
timeout 120 $P --m 16384 --n 4608 --k 64 --cfgs 2x256,2x128 --epi 1
timeout 120 $P --m 16384 --n 4608 --k 64 --cfgs 2x256,2x128 --epi 0
echo "=== K sweep, plain bf16 epilogue N=4608 (mainloop per k-block)"
for k in 576 1152 2304 4608; do timeout 120 $P --m 16384 --n 4608 --k $k --cfgs 2x256; done
echo "=== done"
