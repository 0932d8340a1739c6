#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call13.log 2>&1
echo "=== legacy"
DITB200_ATTN_MMA_SYNC=1 timeout 120 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72
echo "=== tcgen05 small first"
timeout 60 python tools/attn_probe.py --b 1 --t 128 --h 1 --hd 64 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 1 --t 256 --h 1 --hd 64 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 1 --t 256 --h 2 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 2 --t 128 --h 3 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 64; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 8 --t 256 --h 6 --hd 64 --scale 3; echo "exit=$?"
echo "=== done"
