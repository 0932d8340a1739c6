#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c1.log 2>&1
cp ab/base.so fast_dit_b200/lib/libditb200.so
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25
echo "== A/B"
B="--steps 2 --warmup 2"
bash tools/ab_bench.sh r2c1_base_graph ab/base.so -- $B
bash tools/ab_bench.sh r2c1_base_nograph ab/base.so DITB200_GRAPH=0 -- $B
bash tools/ab_bench.sh r2c1_pdl_graph ab/pdl.so -- $B
bash tools/ab_bench.sh r2c1_pdl_nograph ab/pdl.so DITB200_GRAPH=0 -- $B
bash tools/ab_bench.sh r2c1_branch1 ab/base.so DITB200_INFER_BRANCH=1 -- $B
bash tools/ab_bench.sh r2c1_branch2 ab/base.so DITB200_INFER_BRANCH=2 -- $B
bash tools/ab_bench.sh r2c1_mc ab/base.so DITB200_GEMM_MC=1 -- $B
bash tools/ab_bench.sh r2c1_pdl_branch2 ab/pdl.so DITB200_INFER_BRANCH=2 -- $B
bash tools/ab_bench.sh r2c1_base_graph2 ab/base.so -- $B
cp ab/base.so fast_dit_b200/lib/libditb200.so
echo "== sanitizer"; SAN_LIMIT=240 bash tools/sanitize.sh
