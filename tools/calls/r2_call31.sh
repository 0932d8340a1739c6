#!/bin/bash
# Round 2, call 31: adaLN weight gradient through the GEMM route above 64 images (ops.adaln_wgrad_preferred): its
# tests, and the C2 line (DiT-B/4, 256 images per GPU) under the new dispatch.
mkdir -p gpurun_out
exec > gpurun_out/r2c31.log 2>&1
timeout -k 5 40 python -m pytest tests/test_backward_gpu.py -q -m gpu -k "adaln" 2>&1 | tail -3
bash tools/ab_bench.sh r2c31_c2 fast_dit_b200/lib/libditb200.so -- --workload c2 --steps 20 --warmup 5 | head -1
