#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call8.log 2>&1
P="python tools/tc_probe.py --no-cublas"
CMD="$P --m 16384 --n 4608 --k 64 --cfgs 2x256 --iters 2 --sets 1"
timeout 200 $CMD > gpurun_out/plain8.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/r01_gemm_epi_only $CMD > gpurun_out/ncu8.log 2>&1
echo "ncu exit=$?"
