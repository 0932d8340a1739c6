#!/bin/bash
# Round 2, call 24: attention forward with constant-increment descriptors in the MMA issue loops
# (-DDITB200_ATTN_FAST_ISSUE): timeline, same-box A/B of the isolated kernel, attention tests.
mkdir -p gpurun_out
exec > gpurun_out/r2c24.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/default.so
cp ab/fasttrace.so $L
echo "== timeline fasttrace"; timeout -k 10 120 python tools/attn_trace.py > gpurun_out/r2c24_timeline_fasttrace.txt 2>&1; tail -2 gpurun_out/r2c24_timeline_fasttrace.txt
for v in base fast base fast; do
  cp ab/$v.so $L
  echo "== $v"
  timeout -k 10 120 python tools/attn_probe.py --b 64 --t 256 --iters 200 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_probe.py --b 256 --t 128 --iters 100 2>&1 | tail -2
done
cp ab/fast.so $L
echo "== attention tests on fast"; timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" 2>&1 | tail -2
cp /tmp/default.so $L
