#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/call19.log 2>&1
echo "=== pytest gpu"
timeout 1700 python -m pytest tests -q -m gpu --timeout 600 -p no:cacheprovider 2>&1 | grep -v "^$" | tail -12
echo "=== bench c3 (default flags)"
timeout 1200 python bench.py > gpurun_out/bench19_c3.json 2> gpurun_out/bench19_c3.err; echo "bench exit=$?"; tail -3 gpurun_out/bench19_c3.err; cat gpurun_out/bench19_c3.json
echo "=== ncu launch list"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain19.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2060 -c 412 --csv --log-file gpurun_out/r01_launches_v2.csv $CMD > gpurun_out/ncu19.log 2>&1
echo "ncu exit=$?"
echo "=== done"
