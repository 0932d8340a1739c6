#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call42.log 2>&1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -4
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench c3"
timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/bench42_c3.json 2> gpurun_out/bench42_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench42_c3.err; cat gpurun_out/bench42_c3.json
echo "=== bench c4"
timeout 600 python bench.py --workload c4 --steps 5 --warmup 3 > gpurun_out/bench42_c4.json 2> gpurun_out/bench42_c4.err; echo "bench exit=$?"; tail -3 gpurun_out/bench42_c4.err; cat gpurun_out/bench42_c4.json
echo "=== bench c2"
timeout 600 python bench.py --workload c2 --steps 5 --warmup 3 > gpurun_out/bench42_c2.json 2> gpurun_out/bench42_c2.err; echo "bench exit=$?"; tail -3 gpurun_out/bench42_c2.err; cat gpurun_out/bench42_c2.json
echo "=== done"
