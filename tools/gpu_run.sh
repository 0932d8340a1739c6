#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run.log 2>&1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -5
echo "=== bench c4"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_c4.err; cat gpurun_out/bench_c4.json
echo "=== bench c2"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_c2.err; cat gpurun_out/bench_c2.json
