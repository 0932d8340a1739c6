#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call4.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -30
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench4_c3.json 2> gpurun_out/bench4_c3.err; echo "bench exit=$?"; tail -5 gpurun_out/bench4_c3.err; cat gpurun_out/bench4_c3.json
echo "=== bench c4"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench4_c4.json 2> gpurun_out/bench4_c4.err; echo "bench exit=$?"; tail -5 gpurun_out/bench4_c4.err; cat gpurun_out/bench4_c4.json
echo "=== done"
