#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c2.log 2>&1
B="--steps 2 --warmup 2"
bash tools/ab_bench.sh r2c2_pdl_b2 ab/pdl.so DITB200_INFER_BRANCH=2 -- $B
bash tools/ab_bench.sh r2c2_pdl_b2_no192 ab/pdl.so DITB200_INFER_BRANCH=2 DITB200_GEMM_NO192=1 -- $B
bash tools/ab_bench.sh r2c2_pdl_b1_no192 ab/pdl.so DITB200_INFER_BRANCH=1 DITB200_GEMM_NO192=1 -- $B
bash tools/ab_bench.sh r2c2_pdl_b2_no192_ln64 ab/pdl.so DITB200_INFER_BRANCH=2 DITB200_GEMM_NO192=1 DITB200_LN_THREADS=64 -- $B
bash tools/ab_bench.sh r2c2_pdl_b2_no192_ln256 ab/pdl.so DITB200_INFER_BRANCH=2 DITB200_GEMM_NO192=1 DITB200_LN_THREADS=256 -- $B
bash tools/ab_bench.sh r2c2_pdl_b2_nonarrow ab/pdl.so DITB200_INFER_BRANCH=2 DITB200_GEMM_NO192=1 DITB200_NO_NARROW=1 -- $B
bash tools/ab_bench.sh r2c2_pdl_b2_again ab/pdl.so DITB200_INFER_BRANCH=2 -- $B
echo "== new api tests"
cp ab/base.so fast_dit_b200/lib/libditb200.so
timeout 600 python -m pytest -m gpu -q tests/test_diffusion_api_gpu.py tests/test_parity_gpu.py 2>&1 | tail -15
