#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/final.log 2>&1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench reference arm"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit=$?"; cat gpurun_out/bench_ref.json | cut -c1-600
echo "=== bench c3 default"
timeout 1200 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench_c3.err; cat gpurun_out/bench_c3.json
echo "=== done"
