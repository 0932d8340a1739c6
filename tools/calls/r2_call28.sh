#!/bin/bash
# Round 2, call 28: verification pass of the tree with the new attention issue loops and the lazily rescaled
# KV-blocked forward: GPU suite, smoke, default bench line, in-model A/B of C3 and C5 against the previous library
# (ab/base.so), isolated A/B of the KV-blocked kernel (ab/fast.so = eager rescale), C4 line.
mkdir -p gpurun_out
exec > gpurun_out/r2c28.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/new.so
echo "== gpu tests"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== default bench"; timeout -k 10 900 python bench.py > gpurun_out/r2c28_c3_default.json 2> gpurun_out/r2c28_c3_default.err; tail -c 300 gpurun_out/r2c28_c3_default.json; echo
bash tools/ab_bench.sh r2c28_c3_old ab/base.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c3_new /tmp/new.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c5_old ab/base.so -- --workload c5 --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c5_new /tmp/new.so -- --workload c5 --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c4 /tmp/new.so -- --workload c4 --steps 20 --warmup 5 | head -1
for v in fast lazy fast lazy; do
  cp ab/$v.so $L; echo "== kv probe $v"
  timeout -k 10 120 python tools/attn_probe.py --b 16 --t 1024 --iters 50 2>&1 | tail -1
done
cp /tmp/new.so $L
