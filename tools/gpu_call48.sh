#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call48.log 2>&1
CMD="python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72 --iters 3"
timeout 200 $CMD > gpurun_out/plain48.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd_tc -s 2 -c 1 -f -o gpurun_out/r01_attn_bwd_v1 $CMD > gpurun_out/ncu48.log 2>&1
echo "ncu exit=$?"
