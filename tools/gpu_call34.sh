#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call34.log 2>&1
echo "=== T=256: single-pass vs kv kernel"
timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72
DITB200_ATTN_KV=1 timeout 60 python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72
echo "=== pytest attention"
timeout 900 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k attention --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== bench c5"
timeout 900 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench34_c5.json 2> gpurun_out/bench34_c5.err; echo "bench exit=$?"; tail -3 gpurun_out/bench34_c5.err; cat gpurun_out/bench34_c5.json
echo "=== done"
