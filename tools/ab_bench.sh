#!/bin/bash
# Same-box comparison of bench variants: tools/ab_bench.sh <tag> <lib.so> [ENV=VAL ...] -- [bench args]
# Copies <lib.so> over the in-tree library, runs bench.py with the given environment, writes gpurun_out/<tag>.json
# and prints a one-line digest (img/s, per-step breakdown).
tag=$1; lib=$2; shift 2
envs=()
while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done
shift
[ "$lib" -ef fast_dit_b200/lib/libditb200.so ] || cp "$lib" fast_dit_b200/lib/libditb200.so
env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline "$@" > gpurun_out/$tag.json 2> gpurun_out/$tag.err
python - "$tag" <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-1500:]); sys.exit(0)
kb = d.get("kernel_breakdown_ms_per_denoise_step") or d.get("kernel_breakdown_ms_per_step") or {}
top = " ".join(f"{k}={v['ms']:.2f}" for k, v in list(kb.items())[:5])
print(f"{tag}: value={d['value']:.3f} e2e={d['e2e']['value']:.3f} ms/step={d['ms_per_step']:.1f} clk={d['clocks']['sm_mhz']} | {top}")
for k, v in list((d.get("gemm_by_shape") or {}).items())[:6]:
    print(f"    {k}: {v['us_per_launch']:.1f} us x{v['launches']} {v['tflops']:.0f} TF")
P
