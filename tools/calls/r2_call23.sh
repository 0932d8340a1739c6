#!/bin/bash
# Round 2, call 23: timeline of attn_fwd_tc_kernel (clock64 at every hand-off, tools/attn_trace.py) for the current
# kernel and for the deep max pass (-DDITB200_ATTN_DEEP_MAX), and a same-box A/B of the two (isolated kernel time).
mkdir -p gpurun_out
exec > gpurun_out/r2c23.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/default.so
for v in trace deeptrace; do
  cp ab/$v.so $L
  echo "== timeline $v"; timeout -k 10 120 python tools/attn_trace.py > gpurun_out/r2c23_timeline_$v.txt 2>&1; tail -2 gpurun_out/r2c23_timeline_$v.txt
done
for v in base deep base deep; do
  cp ab/$v.so $L
  echo "== $v"
  timeout -k 10 120 python tools/attn_probe.py --b 64 --t 256 --iters 200 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_probe.py --b 256 --t 128 --iters 100 2>&1 | tail -2
done
cp ab/deep.so $L
echo "== attention tests on deep"; timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" 2>&1 | tail -2
cp /tmp/default.so $L
