#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call4.log 2>&1
echo "=== gemm probes (auto tile)"
for shp in "16384 1152 1152 0" "16384 3456 1152 0" "16384 4608 1152 1" "16384 1152 4608 2" "64 195840 1152 0"; do
  set -- $shp
  timeout 120 python tools/tc_probe.py --cg 0 --bn 0 --m $1 --n $2 --k $3 --epi $4 --bench
done
timeout 120 python tools/tc_probe.py --cg 2 --bn 128 --m 16384 --n 1152 --k 1152 --bench
timeout 120 python tools/tc_probe.py --cg 2 --bn 192 --m 16384 --n 1152 --k 1152 --bench
echo "=== pytest gpu (kernels + parity)"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -5
echo "=== bench"
BENCH="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 $BENCH > gpurun_out/bench4.json 2> gpurun_out/bench4.err; rc=$?; echo "bench exit=$rc"; tail -3 gpurun_out/bench4.err; cat gpurun_out/bench4.json
if [ $rc -eq 0 ]; then
  echo "=== ncu launch list (2 denoise steps)"
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 420 --csv --log-file gpurun_out/r01_launches.csv $BENCH > gpurun_out/ncu_launch.log 2>&1
  echo "ncu launches exit=$?"; tail -2 gpurun_out/ncu_launch.log
  echo "=== ncu full on the GEMM (fc1 shape inside the model)"
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 400 -c 4 -o gpurun_out/r01_gemm_full $BENCH > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit=$?"; tail -2 gpurun_out/ncu_full.log
fi
ls -la gpurun_out
echo "=== done"
