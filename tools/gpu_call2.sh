#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call2.log 2>&1
P1="python tools/tc_probe.py --cg 1 --bn 256 --m 16384 --n 4608 --k 1152 --bench"
P2="python tools/tc_probe.py --cg 2 --bn 256 --m 16384 --n 4608 --k 1152 --bench"
$P1 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/gemm_cg1_bn256 $P1
echo "ncu1 exit=$?"
$P2 && ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -o gpurun_out/gemm_cg2_bn256 $P2
echo "ncu2 exit=$?"
ls -la gpurun_out
