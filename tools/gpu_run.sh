#!/bin/bash
mkdir -p gpurun_out
exec > gpurun_out/run2.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 600 python bench.py > gpurun_out/bench_c3_v8.json 2>gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_c3_v8.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])
print(d['kernel_breakdown_ms_per_denoise_step'])
PY
