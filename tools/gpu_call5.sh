#!/bin/bash
# backward bring-up: GEMM variants first (own processes, own timeouts), then the kernel + parity suites
mkdir -p gpurun_out
exec > gpurun_out/call5.log 2>&1
echo "=== gemm forward still fine (8 epilogue warps)"
for shp in "16384 1152 1152 0" "16384 3456 1152 0" "16384 4608 1152 1" "16384 1152 4608 2"; do
  set -- $shp
  timeout 120 python tools/tc_probe.py --cg 0 --bn 0 --m $1 --n $2 --k $3 --epi $4 --bench
done
echo "=== backward tests"
timeout 1500 python -m pytest tests/test_backward_gpu.py -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -40
echo "=== all gpu tests"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | tail -15
echo "=== done"
