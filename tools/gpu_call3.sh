#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call3.log 2>&1
echo "=== cg2 re-probe"
for cfg in "2 128" "2 192" "2 256" "1 256"; do
  set -- $cfg
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 1152 --k 1152 --bench
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 4608 --k 1152 --bench --epi 1
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 3456 --k 1152 --bench
  timeout 120 python tools/tc_probe.py --cg $1 --bn $2 --m 16384 --n 1152 --k 4608 --bench
done
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider -s 2>&1 | grep -v "^$" | tail -60
echo "=== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()"
echo "=== bench (short)"
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench3.json 2> gpurun_out/bench3.err; echo "bench exit=$?"; tail -5 gpurun_out/bench3.err; cat gpurun_out/bench3.json
echo "=== done"
