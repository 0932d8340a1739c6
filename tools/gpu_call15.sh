#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call15.log 2>&1
timeout 60 python tools/attn_probe.py --b 1 --t 128 --h 1 --hd 64 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 1 --t 256 --h 2 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 3 --t 128 --h 5 --hd 72 --iters 3; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 64; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 32 --t 256 --h 16 --hd 72; echo "exit=$?"
timeout 60 python tools/attn_probe.py --b 8 --t 256 --h 6 --hd 64 --scale 3; echo "exit=$?"
echo "=== done"
