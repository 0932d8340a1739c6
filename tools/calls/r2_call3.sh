#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c3.log 2>&1
cp ab/pdl.so fast_dit_b200/lib/libditb200.so
echo "== gpu tests (PDL build, branch mode 2)"; timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
B="--steps 2 --warmup 2"
bash tools/ab_bench.sh r2c3_pdl_b2 ab/pdl.so -- $B
bash tools/ab_bench.sh r2c3_pdl_b2_zig ab/pdl.so DITB200_ZIGZAG=1 -- $B
bash tools/ab_bench.sh r2c3_pdl_b2_2 ab/pdl.so -- $B
bash tools/ab_bench.sh r2c3_pdl_b2_zig_2 ab/pdl.so DITB200_ZIGZAG=1 -- $B
bash tools/ab_bench.sh r2c3_c5_nopdl ab/nopdl.so DITB200_INFER_BRANCH=0 -- --workload c5 $B
bash tools/ab_bench.sh r2c3_c5_pdl ab/pdl.so -- --workload c5 $B
bash tools/ab_bench.sh r2c3_c5_pdl_zig ab/pdl.so DITB200_ZIGZAG=1 -- --workload c5 $B
bash tools/ab_bench.sh r2c3_c4_nopdl ab/nopdl.so -- --workload c4 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c3_c4_pdl ab/pdl.so -- --workload c4 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c3_c2_pdl ab/pdl.so -- --workload c2 --steps 10 --warmup 5
bash tools/ab_bench.sh r2c3_c1_graph ab/pdl.so -- --workload c1 --steps 5 --warmup 3
bash tools/ab_bench.sh r2c3_c1_nograph ab/pdl.so DITB200_GRAPH=0 -- --workload c1 --steps 5 --warmup 3
cp ab/pdl.so fast_dit_b200/lib/libditb200.so
echo "== ncu"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export DITB200_GRAPH=0
$CMD > gpurun_out/r2c3_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 420 --csv --log-file gpurun_out/r2c3_launches.csv $CMD > gpurun_out/r2c3_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ln_modulate -s 60 -c 4 -o gpurun_out/r2c3_ln $CMD > gpurun_out/r2c3_ncu2.log 2>&1
echo "ln full rc=$?"
$CMD > gpurun_out/r2c3_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc_kernel -s 30 -c 2 -o gpurun_out/r2c3_attn $CMD > gpurun_out/r2c3_ncu3.log 2>&1
echo "attn full rc=$?"
ls -la gpurun_out/*.ncu-rep
