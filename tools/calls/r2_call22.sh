#!/bin/bash
# Round 2, call 22: verification pass of the final tree: GPU suite, smoke, default bench line (with the CPU baseline),
# C5 / C4 / C2 lines, ncu launch list of two denoising steps.
mkdir -p gpurun_out
exec > gpurun_out/r2c22.log 2>&1
echo "== gpu tests"; timeout -k 10 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout -k 10 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== default bench"; timeout -k 10 900 python bench.py > gpurun_out/r2c22_c3_default.json 2> gpurun_out/r2c22_c3_default.err; tail -c 400 gpurun_out/r2c22_c3_default.json; echo
L=fast_dit_b200/lib/libditb200.so
bash tools/ab_bench.sh r2c22_c5 $L -- --workload c5 --steps 2 --warmup 2 | head -1
bash tools/ab_bench.sh r2c22_c4 $L -- --workload c4 --steps 20 --warmup 5 | head -1
bash tools/ab_bench.sh r2c22_c2 $L -- --workload c2 --steps 20 --warmup 5 | head -1
echo "== ncu launch list"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export DITB200_GRAPH=0
timeout -k 10 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 420 --csv --log-file gpurun_out/r2c22_launches.csv $CMD > gpurun_out/r2c22_ncu1.log 2>&1
echo "launch list rc=$?"
