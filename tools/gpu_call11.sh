#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call11.log 2>&1
P="python tools/tc_probe.py"
echo "=== correctness (narrow tiles)"
timeout 120 $P --m 4096 --n 1152 --k 1152 --cfgs 2x256,2x192,2x128,1x256,0x0 --check --iters 5 --no-cublas
timeout 120 $P --m 3840 --n 1000 --k 1152 --cfgs 2x256,1x192,0x0 --check --iters 5 --no-cublas --epi 2
timeout 120 $P --m 4096 --n 1000 --k 1152 --cfgs 2x256,0x0 --check --iters 5 --no-cublas --epi 1
timeout 120 $P --m 4096 --n 1152 --k 1152 --cfgs 2x256,0x0 --check --iters 5 --no-cublas --trans-w
timeout 120 $P --m 1152 --n 1152 --k 4096 --cfgs 2x256,0x0 --check --iters 5 --no-cublas --trans-w --trans-a --split-k 4
echo "=== C3 shapes"
timeout 120 $P --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x192,0x0 --epi 2
timeout 120 $P --m 16384 --n 1152 --k 4608 --cfgs 2x256,2x192,0x0 --epi 2
timeout 120 $P --m 16384 --n 3456 --k 1152 --cfgs 2x256,2x192,0x0
timeout 120 $P --m 16384 --n 4608 --k 1152 --cfgs 2x256,2x192,0x0 --epi 1
echo "=== C4 shapes (M=8192)"
timeout 120 $P --m 8192 --n 1152 --k 1152 --cfgs 2x256,2x192,0x0 --epi 2
timeout 120 $P --m 8192 --n 1152 --k 4608 --cfgs 2x256,2x192,0x0 --epi 2
timeout 120 $P --m 8192 --n 3456 --k 1152 --cfgs 2x256,2x192,0x0
timeout 120 $P --m 8192 --n 4608 --k 1152 --cfgs 2x256,2x192,0x0 --epi 1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -8
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench11_c3.json 2> gpurun_out/bench11_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench11_c3.err; cat gpurun_out/bench11_c3.json
echo "=== done"
