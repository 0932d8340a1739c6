#!/bin/bash
# Round 2, call 29: ncu --set full of the two attention forward kernels after the third session's changes
# (isolated launches of tools/attn_probe.py: T = 256 at C3's shape, T = 1024 at C5's).
mkdir -p gpurun_out
exec > gpurun_out/r2c29.log 2>&1
timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kernel -s 3 -c 1 -o gpurun_out/r2c29_attn256 python tools/attn_probe.py --b 64 --t 256 --iters 3 2>&1 | tail -3
timeout -k 10 200 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kv_kernel -s 3 -c 1 -o gpurun_out/r2c29_attn1024 python tools/attn_probe.py --b 16 --t 1024 --iters 3 2>&1 | tail -3
ls -la gpurun_out/r2c29_*
