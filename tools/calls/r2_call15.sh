#!/bin/bash
# Round 2, call 15: fused LayerNorm-backward + gate backward, adaLN outer-product weight gradient, batched d silu(c):
# kernel tests, model-gradient tests, same-box A/B of the training step against the previous commit (ab/base_tree).
mkdir -p gpurun_out
exec > gpurun_out/r2c15.log 2>&1
ROOT=$PWD
echo "== backward tests"; timeout 900 python -m pytest tests/test_backward_gpu.py -m gpu -x -q 2>&1 | tail -6
run() { # tag, dir, args
  tag=$1; dir=$2; shift 2
  (cd $dir && timeout 600 python bench.py --no-cpu-baseline "$@" > $ROOT/gpurun_out/$tag.json 2> $ROOT/gpurun_out/$tag.err)
  python - $tag <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    kb = d.get("kernel_breakdown_ms_per_step", {})
    top = " ".join(f"{k}={v['ms']:.2f}" for k, v in list(kb.items())[:12])
    print(f"{tag}: value={d['value']:.1f} ms/step={d['ms_per_step']:.2f} e2e={d['e2e']['value']:.1f} clk={d['clocks']['sm_mhz']} W={d['clocks'].get('power_w_max')} | {top}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-1500:])
P
}
A="--workload c4 --steps 20 --warmup 5"
run r2c15_c4_base ab/base_tree $A
run r2c15_c4_new . $A
run r2c15_c4_base_b ab/base_tree $A
run r2c15_c4_new_b . $A
A="--workload c2 --steps 20 --warmup 5"
run r2c15_c2_base ab/base_tree $A
run r2c15_c2_new . $A
echo "== per-shape profile (new)"
timeout 300 python tools/train_profile.py --workload c4 2>&1 | tail -48
