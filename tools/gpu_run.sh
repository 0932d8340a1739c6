#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run2.log 2>&1
DITB200_ATTN_DIRECT_OUT=1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" 2>&1 | tail -2
for rep in 1 2; do
echo "== staged"; timeout 120 python tools/attn_probe.py --iters 100 2>&1 | tail -1
echo "== direct"; DITB200_ATTN_DIRECT_OUT=1 timeout 120 python tools/attn_probe.py --iters 100 2>&1 | tail -1
done
echo "== staged T128"; timeout 120 python tools/attn_probe.py --iters 100 --t 128 --b 256 2>&1 | tail -1
echo "== direct T128"; DITB200_ATTN_DIRECT_OUT=1 timeout 120 python tools/attn_probe.py --iters 100 --t 128 --b 256 2>&1 | tail -1
