#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call44.log 2>&1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 300 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -6
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench c3"
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench44_c3.json 2> gpurun_out/bench44_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench44_c3.err; cat gpurun_out/bench44_c3.json
echo "=== done"
