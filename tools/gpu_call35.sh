#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call35.log 2>&1
P="python tools/tc_probe.py --no-cublas"
echo "=== TMA store correctness"
timeout 60 $P --m 4096 --n 1152 --k 1152 --cfgs 0x0,2x256,2x128,1x256 --check --iters 3; echo "exit=$?"
timeout 60 $P --m 4000 --n 1000 --k 1152 --cfgs 0x0,1x192 --check --iters 3 --epi 1; echo "exit=$?"
timeout 60 $P --m 300 --n 200 --k 72 --cfgs 0x0 --check --iters 3; echo "exit=$?"
echo "=== timing (TMA store)"
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1
timeout 90 $P --m 16384 --n 3456 --k 64 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 64 --cfgs 0x0 --epi 1
timeout 90 $P --m 8192 --n 3456 --k 1152 --cfgs 0x0
echo "=== timing (no TMA store)"
export DITB200_NO_TMA_STORE=1
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1
timeout 90 $P --m 16384 --n 3456 --k 64 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 64 --cfgs 0x0 --epi 1
timeout 90 $P --m 8192 --n 3456 --k 1152 --cfgs 0x0
echo "=== done"
