#!/bin/bash
# Round 2, call 27: attention forward, exp pass software-pipelined by 16 keys (-DDITB200_ATTN_PIPE) against the
# current kernel: timeline, same-box A/B, attention tests.
mkdir -p gpurun_out
exec > gpurun_out/r2c27.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/default.so
cp ab/pipetrace.so $L
echo "== timeline pipetrace"; timeout -k 10 120 python tools/attn_trace.py > gpurun_out/r2c27_timeline_pipetrace.txt 2>&1; tail -2 gpurun_out/r2c27_timeline_pipetrace.txt
for v in fast pipe fast pipe; do
  cp ab/$v.so $L
  echo "== $v"
  timeout -k 10 120 python tools/attn_probe.py --b 64 --t 256 --iters 200 2>&1 | tail -2
  timeout -k 10 120 python tools/attn_probe.py --b 256 --t 128 --iters 100 2>&1 | tail -2
done
cp ab/pipe.so $L
echo "== attention tests on pipe"; timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -q -k "attention" 2>&1 | tail -2
cp /tmp/default.so $L
