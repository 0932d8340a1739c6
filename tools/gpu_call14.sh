#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call14.log 2>&1
CMD="python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72 --iters 3"
timeout 200 $CMD > gpurun_out/plain14.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 3 -c 1 -f -o gpurun_out/r01_attn_tc_v2 $CMD > gpurun_out/ncu14.log 2>&1
echo "ncu exit=$?"
