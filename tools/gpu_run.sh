#!/bin/bash
# scratch driver for one gpurun call: edit, run with `gpurun -- bash tools/gpu_run.sh`, read gpurun_out/run.log
mkdir -p gpurun_out
exec > gpurun_out/run.log 2>&1
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72
timeout 60 python tools/attn_probe.py --b 16 --t 1024 --h 16 --hd 72 --iters 10
timeout 60 python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72
echo "=== pytest attention"
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -q -m gpu -k attention --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -3
