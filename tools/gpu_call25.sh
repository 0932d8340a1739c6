#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call25.log 2>&1
P="python tools/tc_probe.py --no-cublas"
timeout 90 $P --m 8192 --n 4608 --k 64 --cfgs 0x0 --trans-w --epi 4
timeout 90 $P --m 8192 --n 4608 --k 64 --cfgs 0x0 --trans-w --epi 0
timeout 90 $P --m 8192 --n 4608 --k 64 --cfgs 0x0 --epi 1
CMD="$P --m 8192 --n 4608 --k 1152 --cfgs 0x0 --trans-w --epi 4 --iters 2 --sets 1"
timeout 200 $CMD > gpurun_out/plain25.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/r01_gemm_dgelu $CMD > gpurun_out/ncu25.log 2>&1
echo "ncu exit=$?"
