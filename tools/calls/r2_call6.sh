#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c6.log 2>&1
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
bash tools/ab_bench.sh r2c6_c4_overlap fast_dit_b200/lib/libditb200.so -- --workload c4 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c6_c4_nooverlap fast_dit_b200/lib/libditb200.so -- --workload c4 --steps 10 --warmup 5 --no-overlap-opt
bash tools/ab_bench.sh r2c6_c4_overlap2 fast_dit_b200/lib/libditb200.so -- --workload c4 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c6_c2_overlap fast_dit_b200/lib/libditb200.so -- --workload c2 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c6_c2_nooverlap fast_dit_b200/lib/libditb200.so -- --workload c2 --steps 10 --warmup 5 --no-overlap-opt
