#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call33.log 2>&1
echo "=== legacy T=1024"
DITB200_ATTN_MMA_SYNC=1 timeout 120 python tools/attn_probe.py --b 16 --t 1024 --h 16 --hd 72 --iters 10
echo "=== tcgen05 kv"
timeout 60 python tools/attn_probe.py --b 1 --t 512 --h 1 --hd 64 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 1 --t 1024 --h 2 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 2 --t 768 --h 3 --hd 72 --iters 3 --scale 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 16 --t 1024 --h 16 --hd 72 --iters 10; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 16 --t 1024 --h 16 --hd 64 --iters 10; echo "exit=$?"
echo "=== done"
