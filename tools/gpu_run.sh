#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run2.log 2>&1
L=fast_dit_b200/lib/libditb200.so
for v in cur k192 cur k192; do
  cp ab/libditb200_$v.so $L
  echo "== $v"
  timeout 300 python bench.py --workload c4 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('   ', round(d['value'],1),'img/s', round(d['ms_per_step'],2),'ms/step')"
done
cp ab/libditb200_cur.so $L
