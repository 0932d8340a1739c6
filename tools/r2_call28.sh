#!/bin/bash
# Round 2, call 28: verification pass of the tree with the new attention issue loops: GPU suite, smoke, default bench
# line, in-model A/B of C3 against the previous library (ab/base.so), C5 and C4 lines.
mkdir -p gpurun_out
exec > gpurun_out/r2c28.log 2>&1
L=fast_dit_b200/lib/libditb200.so
cp $L /tmp/new.so
echo "== gpu tests"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== default bench"; timeout -k 10 900 python bench.py > gpurun_out/r2c28_c3_default.json 2> gpurun_out/r2c28_c3_default.err; tail -c 300 gpurun_out/r2c28_c3_default.json; echo
bash tools/ab_bench.sh r2c28_c3_old ab/base.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c3_new /tmp/new.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c3_old_b ab/base.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c3_new_b /tmp/new.so -- --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c28_c4 /tmp/new.so -- --workload c4 --steps 20 --warmup 5 | head -1
cp /tmp/new.so $L
