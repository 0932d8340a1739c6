#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c14.log 2>&1
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -5
B="--steps 2 --warmup 2"
L=fast_dit_b200/lib/libditb200.so
bash tools/ab_bench.sh r2c14_mode3 $L -- $B
bash tools/ab_bench.sh r2c14_mode2 $L DITB200_INFER_BRANCH=2 -- $B
bash tools/ab_bench.sh r2c14_mode3b $L -- $B
bash tools/ab_bench.sh r2c14_mode2b $L DITB200_INFER_BRANCH=2 -- $B
bash tools/ab_bench.sh r2c14_c5_mode3 $L -- --workload c5 $B
bash tools/ab_bench.sh r2c14_c5_mode2 $L DITB200_INFER_BRANCH=2 -- --workload c5 $B
