#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call24.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== dgelu epilogue"
timeout 90 $P --m 8192 --n 4608 --k 1152 --cfgs 0x0 --trans-w --epi 4 --check
timeout 90 $P --m 8192 --n 4608 --k 1152 --cfgs 0x0 --trans-w --epi 0
echo "=== wgrad: layout effect (same flops; K-major vs MN-major operands), split 4 and 1"
for sk in 4 1; do
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 2x256 --split-k $sk
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 2x256 --split-k $sk --trans-a
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 2x256 --split-k $sk --trans-w
timeout 90 $P --m 1152 --n 4608 --k 8192 --cfgs 2x256 --split-k $sk --trans-a --trans-w
done
echo "=== done"
