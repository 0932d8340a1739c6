#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c7.log 2>&1
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
bash tools/ab_bench.sh r2c7_c4 fast_dit_b200/lib/libditb200.so -- --workload c4 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c7_c4_overlap fast_dit_b200/lib/libditb200.so -- --workload c4 --steps 10 --warmup 5 --overlap-opt
bash tools/ab_bench.sh r2c7_c2 fast_dit_b200/lib/libditb200.so -- --workload c2 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c7_c5 fast_dit_b200/lib/libditb200.so -- --workload c5 --steps 2 --warmup 2
bash tools/ab_bench.sh r2c7_c3 fast_dit_b200/lib/libditb200.so -- --steps 2 --warmup 2
