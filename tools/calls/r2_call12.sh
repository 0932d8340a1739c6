#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/r2c12.log 2>&1
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4
B="--steps 2 --warmup 2"
L=fast_dit_b200/lib/libditb200.so
bash tools/ab_bench.sh r2c12_c3 $L -- $B
bash tools/ab_bench.sh r2c12_c3_notable $L DITB200_GEMM_NO_TABLE=1 -- $B
bash tools/ab_bench.sh r2c12_c3_b $L -- $B
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== ncu"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
export DITB200_GRAPH=0
$CMD > gpurun_out/r2c12_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 420 --csv --log-file gpurun_out/r2c12_launches.csv $CMD > gpurun_out/r2c12_ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r2c12_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 42 -c 4 -o gpurun_out/r2c12_gemm $CMD > gpurun_out/r2c12_ncu2.log 2>&1
echo "gemm full rc=$?"
ls -la gpurun_out/r2c12_gemm.ncu-rep
