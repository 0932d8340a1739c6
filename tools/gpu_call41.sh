#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call41.log 2>&1
P="python tools/tc_probe.py --no-cublas"
python tools/dbg_resid.py 2>&1 | grep "err"
echo "=== correctness"
timeout 60 $P --m 4096 --n 1152 --k 1152 --cfgs 0x0,2x256,2x192,2x128,1x256 --check --iters 3 --epi 2 --inplace; echo "exit=$?"
timeout 60 $P --m 3840 --n 1000 --k 1152 --cfgs 0x0,1x192 --check --iters 3 --epi 2; echo "exit=$?"
echo "=== timing"
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 0x0,2x256 --epi 2 --inplace
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 0x0,2x256 --epi 2 --inplace
timeout 90 $P --m 16384 --n 1152 --k 64 --cfgs 0x0 --epi 2 --inplace
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench41_c3.json 2> gpurun_out/bench41_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench41_c3.err; cat gpurun_out/bench41_c3.json
echo "=== done"
