#!/bin/bash
# Round 2, call 17: attention forward with one MMA-issuing warp per query tile (T <= 256 kernel and the KV-blocked
# kernel): isolated probe new vs base, GPU suite, same-box C3 / C5 A/B.
mkdir -p gpurun_out
exec > gpurun_out/r2c17.log 2>&1
LIB=fast_dit_b200/lib/libditb200.so
probe() { timeout -k 10 120 python tools/attn_probe.py "$@" 2>&1 | tail -2; }
for v in new base new base; do
  cp ab/$v.so $LIB
  echo "== probe $v"
  probe --b 64 --t 256 --h 16 --hd 72
  probe --b 256 --t 128 --h 12 --hd 64
  probe --b 16 --t 1024 --h 16 --hd 72
  probe --b 32 --t 256 --h 16 --hd 72
done
cp ab/new.so $LIB
echo "== gpu tests (new)"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
B="--steps 1 --warmup 1"
bash tools/ab_bench.sh r2c17_c3_base ab/base.so -- $B
bash tools/ab_bench.sh r2c17_c3_new ab/new.so -- $B
bash tools/ab_bench.sh r2c17_c3_base_b ab/base.so -- $B
bash tools/ab_bench.sh r2c17_c3_new_b ab/new.so -- $B
bash tools/ab_bench.sh r2c17_c5_base ab/base.so -- --workload c5 $B
bash tools/ab_bench.sh r2c17_c5_new ab/new.so -- --workload c5 $B
cp ab/new.so $LIB
