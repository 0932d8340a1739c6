#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call47.log 2>&1
echo "=== legacy"
DITB200_ATTN_MMA_SYNC=1 timeout 120 python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72
echo "=== tcgen05"
timeout 60 python tools/attn_bwd_probe.py --b 1 --t 256 --h 1 --hd 64 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_bwd_probe.py --b 1 --t 256 --h 2 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72; echo "exit=$?"
timeout 60 python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 64; echo "exit=$?"
echo "=== done"
