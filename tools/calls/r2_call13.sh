#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c13.log 2>&1
B="--steps 2 --warmup 2"
L=fast_dit_b200/lib/libditb200.so
bash tools/ab_bench.sh r2c13_base $L -- $B
bash tools/ab_bench.sh r2c13_defer_timing_only $L DITB200_EXP_DEFER=1 -- $B
bash tools/ab_bench.sh r2c13_base2 $L -- $B
bash tools/ab_bench.sh r2c13_defer_timing_only2 $L DITB200_EXP_DEFER=1 -- $B
