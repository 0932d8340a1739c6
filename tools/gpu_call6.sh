#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call6.log 2>&1
echo "=== backward tests"
timeout 1500 python -m pytest tests/test_backward_gpu.py -q -x --timeout 300 -p no:cacheprovider 2>&1 | tail -30
echo "=== train bench c4 (XL/2, 32 img)"
timeout 900 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/train_c4.json 2> gpurun_out/train_c4.err; echo "rc=$?"; tail -5 gpurun_out/train_c4.err; cat gpurun_out/train_c4.json
echo "=== train bench c2 (B/4, 256 img)"
timeout 900 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/train_c2.json 2> gpurun_out/train_c2.err; echo "rc=$?"; tail -5 gpurun_out/train_c2.err; cat gpurun_out/train_c2.json
echo "=== done"
