#!/bin/bash
# Round 2, call 25: attention forward, exp passes of the two query tiles serialised by a per-sub-partition token
# (-DDITB200_ATTN_XU_TOKEN on top of the fast issue loops): timeline, same-box A/B, attention tests.
mkdir -p gpurun_out
exec > gpurun_out/r2c25.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/default.so
cp ab/toktrace.so $L
echo "== timeline toktrace"; timeout -k 10 120 python tools/attn_trace.py > gpurun_out/r2c25_timeline_toktrace.txt 2>&1; tail -2 gpurun_out/r2c25_timeline_toktrace.txt
for v in fast tok fast tok; do
  cp ab/$v.so $L
  echo "== $v"
  timeout -k 10 120 python tools/attn_probe.py --b 64 --t 256 --iters 200 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_probe.py --b 256 --t 128 --iters 100 2>&1 | tail -2
done
cp ab/tok.so $L
echo "== attention tests on tok"; timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" 2>&1 | tail -2
cp /tmp/default.so $L
