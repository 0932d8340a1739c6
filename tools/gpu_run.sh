#!/bin/bash
# Standard verification pass on a GPU box (run through gpurun from the repo root):
#   gpurun --timeout 1800 -- 'bash tools/gpu_run.sh'
# GPU parity tests, the smoke entry point, then the default bench line; everything lands in gpurun_out/.
# Every step runs under its own timeout so that a hung kernel cannot hold the box.
mkdir -p gpurun_out
exec > gpurun_out/verify.log 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_c3.json 2> gpurun_out/bench_err.log
tail -c 600 gpurun_out/bench_c3.json
