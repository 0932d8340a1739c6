#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c4.log 2>&1
echo "== attention probes (new 2-threads-per-row kernel vs round-1 kernel)"
for cfg in "64 256 16 72" "256 128 12 64" "64 256 6 64" "11 256 16 80"; do
  set -- $cfg
  timeout 120 python tools/attn_probe.py --b $1 --t $2 --h $3 --hd $4 --scale 2.0 2>&1 | tail -2
  DITB200_ATTN_1T=1 timeout 120 python tools/attn_probe.py --b $1 --t $2 --h $3 --hd $4 --scale 2.0 2>&1 | tail -1
done
echo "== gpu tests"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
B="--steps 2 --warmup 2"
bash tools/ab_bench.sh r2c4_new fast_dit_b200/lib/libditb200.so -- $B
bash tools/ab_bench.sh r2c4_attn1t fast_dit_b200/lib/libditb200.so DITB200_ATTN_1T=1 -- $B
bash tools/ab_bench.sh r2c4_new2 fast_dit_b200/lib/libditb200.so -- $B
bash tools/ab_bench.sh r2c4_attn1t2 fast_dit_b200/lib/libditb200.so DITB200_ATTN_1T=1 -- $B
