#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call28.log 2>&1
P="python tools/tc_probe.py --no-cublas"
timeout 90 $P --m 8192 --n 4608 --k 1152 --cfgs 0x0 --trans-w --epi 4 --check
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 0x0 --epi 2 --inplace --check
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 0x0 --epi 2 --inplace
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 0x0 --check
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1
timeout 90 $P --m 8192 --n 1152 --k 1152 --cfgs 0x0 --epi 2
