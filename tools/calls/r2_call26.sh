#!/bin/bash
# Round 2, call 26: (a) constant-increment descriptors also in the KV-blocked forward and the backward kernel
# (fast = T <= 256 forward only, fast2 = all three); (b) exp pass software-pipelined by one chunk with integer-pipe
# bf16 rounding (-DDITB200_ATTN_EXP2) against fast2; timeline of (b); attention tests on exp2.
mkdir -p gpurun_out
exec > gpurun_out/r2c26.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/default.so
cp ab/exp2trace.so $L
echo "== timeline exp2trace"; timeout -k 10 120 python tools/attn_trace.py > gpurun_out/r2c26_timeline_exp2trace.txt 2>&1; tail -2 gpurun_out/r2c26_timeline_exp2trace.txt
for v in fast fast2 exp2 fast fast2 exp2; do
  cp ab/$v.so $L
  echo "== $v"
  timeout -k 10 120 python tools/attn_probe.py --b 64 --t 256 --iters 200 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_probe.py --b 16 --t 1024 --iters 50 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_bwd_probe.py --b 32 --t 256 --iters 50 2>&1 | tail -2
done
cp ab/exp2.so $L
echo "== attention tests on exp2"; timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py tests/test_backward_gpu.py -q -k "attention" 2>&1 | tail -2
cp /tmp/default.so $L
