#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call46.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== ncu full: GEMM fc1 (GELU, TMA store) + qkv"
CMD="$P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1 --iters 2 --sets 1"
timeout 200 $CMD > gpurun_out/plain46.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 5 -c 1 -f -o gpurun_out/r01_gemm_fc1_v2 $CMD > gpurun_out/ncu46a.log 2>&1
echo "ncu exit=$?"
CMD2="python tools/attn_probe.py --b 16 --t 1024 --h 16 --hd 72 --iters 3"
timeout 200 $CMD2 > gpurun_out/plain46b.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kv -s 3 -c 1 -f -o gpurun_out/r01_attn_kv_v1 $CMD2 > gpurun_out/ncu46b.log 2>&1
echo "ncu exit=$?"
echo "=== done"
