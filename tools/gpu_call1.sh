#!/bin/bash
# First GPU bring-up: probe the tcgen05 GEMM config by config (each under its own timeout),
# then the per-kernel pytest suite.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
exec > gpurun_out/call1.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv
python -c "import torch; print(torch.__version__, torch.cuda.get_device_name(0))"
for cfg in "1 128 256 256 64" "1 128 256 256 256" "1 192 256 384 384" "1 256 512 512 512" \
           "2 128 256 256 64" "2 128 512 256 256" "2 192 512 384 384" "2 256 512 512 512"; do
  set -- $cfg
  echo "=== probe cg=$1 bn=$2 M=$3 N=$4 K=$5"
  timeout 90 python tools/tc_probe.py --cg $1 --bn $2 --m $3 --n $4 --k $5
  echo "exit=$?"
done
echo "=== structured pattern probe"
timeout 90 python tools/tc_probe.py --cg 1 --bn 128 --m 256 --n 256 --k 128 --pattern struct
timeout 90 python tools/tc_probe.py --cg 2 --bn 128 --m 256 --n 256 --k 128 --pattern struct
echo "=== big shapes + bench"
for cfg in "1 128" "1 192" "1 256" "2 128" "2 192" "2 256"; do
  set -- $cfg
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 1152 --k 1152 --bench
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 4608 --k 1152 --bench --epi 1
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 1152 --k 4608 --bench
done
echo "=== pytest kernels"
timeout 1500 python -m pytest tests/test_kernels_gpu.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -80
echo "=== done"
