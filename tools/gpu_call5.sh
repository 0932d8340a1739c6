#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call5.log 2>&1
P="python tools/tc_probe.py"
echo "=== correctness of new tile widths"
timeout 120 $P --m 4096 --n 1152 --k 1152 --cfgs 2x144,2x256,1x144 --check --iters 5 --no-cublas
timeout 120 $P --m 4096 --n 1152 --k 1152 --cfgs 2x144 --check --iters 5 --no-cublas --epi 2
timeout 120 $P --m 4096 --n 1000 --k 1152 --cfgs 2x144 --check --iters 5 --no-cublas --epi 1
echo "=== C3 shapes"
timeout 120 $P --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x192,2x144,2x128 --epi 2
timeout 120 $P --m 16384 --n 1152 --k 4608 --cfgs 2x256,2x192,2x144,2x128 --epi 2
timeout 120 $P --m 16384 --n 3456 --k 1152 --cfgs 2x256,2x192,2x144,2x128
timeout 120 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,2x192,2x144,2x128 --epi 1
echo "=== perfectly quantised big shapes (74 pairs): intensity test"
timeout 120 $P --m 18944 --n 4608 --k 4608 --cfgs 2x256,2x192,2x144,2x128
echo "=== ncu"
CMD="$P --m 16384 --n 3456 --k 1152 --cfgs 2x256,2x144 --iters 2 --sets 1 --no-cublas"
timeout 200 $CMD > gpurun_out/plain5.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 8 -c 2 -f -o gpurun_out/r01_gemm_qkv $CMD > gpurun_out/ncu5.log 2>&1
echo "ncu exit=$?"
echo "=== done"
