#!/bin/bash
# Round 2, call 16: full GPU suite on the fused-backward tree; adaLN weight-gradient kernel with asynchronous tile
# copies; persistent-launch switch of the LayerNorm forward kernels (C3 A/B); ncu captures of the new kernels.
mkdir -p gpurun_out
exec > gpurun_out/r2c16.log 2>&1
ROOT=$PWD
echo "== gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run() { # tag, dir, env..., -- args
  tag=$1; dir=$2; shift 2
  envs=()
  while [ "$1" != "--" ] && [ $# -gt 0 ]; do envs+=("$1"); shift; done
  shift
  (cd $dir && env "${envs[@]}" timeout 600 python bench.py --no-cpu-baseline "$@" > $ROOT/gpurun_out/$tag.json 2> $ROOT/gpurun_out/$tag.err)
  python - $tag <<'P'
import json, sys
tag = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/{tag}.json").read().strip().splitlines()[-1])
    kb = d.get("kernel_breakdown_ms_per_step") or d.get("kernel_breakdown_ms_per_denoise_step") or {}
    top = " ".join(f"{k}={v['ms']:.2f}" for k, v in list(kb.items())[:10])
    print(f"{tag}: value={d['value']:.3f} ms/step={d['ms_per_step']:.2f} e2e={d['e2e']['value']:.3f} clk={d['clocks']['sm_mhz']} W={d['clocks'].get('power_w_max')} | {top}")
except Exception as e:
    print(tag, "FAILED", e); print(open(f"gpurun_out/{tag}.err").read()[-1500:])
P
}
A="--workload c4 --steps 20 --warmup 5"
run r2c16_c4_new . -- $A
run r2c16_c4_new_b . -- $A
A="--workload c3 --steps 1 --warmup 1"
run r2c16_c3_ln0 . -- $A
run r2c16_c3_ln6 . DITB200_LN_CTAS_PER_SM=6 -- $A
run r2c16_c3_ln7 . DITB200_LN_CTAS_PER_SM=7 -- $A
run r2c16_c3_ln0_b . -- $A
run r2c16_c3_ln6_b . DITB200_LN_CTAS_PER_SM=6 -- $A
run r2c16_c3_ln7_b . DITB200_LN_CTAS_PER_SM=7 -- $A
echo "== C4 with the persistent LayerNorm forward"
A="--workload c4 --steps 20 --warmup 5"
run r2c16_c4_ln7 . DITB200_LN_CTAS_PER_SM=7 -- $A
echo "== ncu: fused LayerNorm backward, adaLN weight gradient"
timeout 300 python tools/train_profile.py --workload c4 > gpurun_out/r2c16_tp.log 2>&1 && tail -n +1 gpurun_out/r2c16_tp.log | head -24
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ln_modulate_bwd_cols_kernel|adaln_wgrad_kernel" -s 20 -c 4 -o gpurun_out/r2c16_bwd python tools/train_profile.py --workload c4 > gpurun_out/r2c16_ncu.log 2>&1
echo "ncu rc=$?"; ls -la gpurun_out/r2c16_bwd.ncu-rep
