#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call49.log 2>&1
timeout 60 python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72; echo "exit=$?"
timeout 60 python tools/attn_bwd_probe.py --b 37 --t 256 --h 16 --hd 72; echo "exit=$?"
echo "=== pytest backward"
timeout 900 python -m pytest tests/test_backward_gpu.py -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c4"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench49_c4.json 2> gpurun_out/bench49_c4.err; echo "bench exit=$?"; tail -3 gpurun_out/bench49_c4.err; cat gpurun_out/bench49_c4.json
