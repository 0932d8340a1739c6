#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call37.log 2>&1
P="python tools/tc_probe.py --no-cublas"
timeout 60 $P --m 4096 --n 1152 --k 1152 --cfgs 2x192,1x192,2x256 --check --iters 3; echo "exit=$?"
timeout 60 $P --m 4000 --n 1000 --k 1152 --cfgs 0x0,1x192,2x192,2x128 --check --iters 3 --epi 1; echo "exit=$?"
timeout 60 $P --m 4000 --n 1096 --k 1152 --cfgs 2x256,2x192,2x128 --check --iters 3 --epi 1; echo "exit=$?"
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench37_c3.json 2> gpurun_out/bench37_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench37_c3.err; cat gpurun_out/bench37_c3.json
echo "=== done"
