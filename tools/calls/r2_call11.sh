#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c11.log 2>&1
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6
B="--steps 2 --warmup 2"
L=fast_dit_b200/lib/libditb200.so
bash tools/ab_bench.sh r2c11_table $L -- $B
bash tools/ab_bench.sh r2c11_notable $L DITB200_GEMM_NO_TABLE=1 -- $B
bash tools/ab_bench.sh r2c11_table2 $L -- $B
bash tools/ab_bench.sh r2c11_notable2 $L DITB200_GEMM_NO_TABLE=1 -- $B
bash tools/ab_bench.sh r2c11_c5_table $L -- --workload c5 $B
bash tools/ab_bench.sh r2c11_c5_notable $L DITB200_GEMM_NO_TABLE=1 -- --workload c5 $B
