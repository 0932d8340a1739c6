#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call27.log 2>&1
P="python tools/tc_probe.py --no-cublas"
timeout 90 $P --m 8192 --n 4608 --k 64 --cfgs 0x0 --trans-w --epi 4
timeout 90 $P --m 8192 --n 4608 --k 1152 --cfgs 0x0 --trans-w --epi 4 --check
timeout 90 $P --m 16384 --n 1152 --k 1152 --cfgs 0x0 --epi 2 --inplace --check
timeout 90 $P --m 16384 --n 1152 --k 4608 --cfgs 0x0 --epi 2 --inplace
timeout 90 $P --m 16384 --n 3456 --k 1152 --cfgs 0x0
timeout 90 $P --m 16384 --n 4608 --k 1152 --cfgs 0x0 --epi 1
timeout 90 $P --m 8192 --n 1152 --k 1152 --cfgs 0x0 --epi 2
timeout 90 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench27_c3.json 2> gpurun_out/bench27_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench27_c3.err; cat gpurun_out/bench27_c3.json
echo "=== bench c4"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench27_c4.json 2> gpurun_out/bench27_c4.err; echo "bench exit=$?"; tail -3 gpurun_out/bench27_c4.err; cat gpurun_out/bench27_c4.json
echo "=== done"
